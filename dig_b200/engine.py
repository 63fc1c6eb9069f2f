"""Drop-in boundary B2: `train_one_epoch` with the reference's signature and return value
(engine_for_pretraining_moco.py:26-204, tag E), re-sequenced for the B200 path:

  * target build + masked gather + MSE (E:83-111, E:141) is ONE kernel (dig_masked_mse) that patchifies
    the un-normalised view-0 pixels on the fly;
  * the model call is one autograd node running the sm_100a kernels (no autocast: operands are bf16 by
    construction, statistics / losses fp32);
  * the ~10 blocking `.item()` reads per step (E:123-176) are replaced by ONE packed device->host copy, consumed one step late;
    a non-finite loss still never reaches the weights: the optimizer launch is guarded by the loss value on the device.
"""
import math
import os
import sys

import numpy as np
import torch

from . import ops, utils
from .ops import call


class _MaskedPixelMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, images, idx, normalize_target=False):
        n_rows = idx.numel()
        p = pred.reshape(n_rows, 48).contiguous()
        loss = torch.zeros(1, dtype=torch.float32, device=pred.device)
        dpred = torch.empty_like(p)
        call("dig_masked_mse", p, images, idx, loss, dpred, n_rows, int(bool(normalize_target)))
        ctx.save_for_backward(dpred)
        ctx.shape = pred.shape
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        out = torch.empty_like(dpred)
        call("dig_scale_by_device_scalar", dpred, g.reshape(1).to(torch.float32).contiguous(), out, dpred.numel())
        return out.view(ctx.shape), None, None, None


def masked_pixel_mse(pred, images, mask_view0, normalize_target=False):
    """F.mse_loss(pred, patchify(unnormalise(images))[mask]) (E:85-111,141).  pred fp32 [B,n,48]; images fp32 [B,3,32,128]
    normalised with mean=std=0.5; mask_view0 bool/uint8 [B,256] with exactly n set positions per sample.  normalize_target: the
    per-patch standardised target of E:89-94 (`normlize_target=True`)."""
    B, n = pred.shape[0], pred.shape[1]
    if not pred.is_cuda:
        raise ops.DigError("masked_pixel_mse runs on CUDA tensors only")
    idx = torch.empty(B * n, dtype=torch.int32, device=pred.device)
    err = torch.zeros(1, dtype=torch.int32, device=pred.device)
    call("dig_mask_to_index", mask_view0.to(torch.uint8).contiguous(), idx, err, B, n)      # always writes B*n in-bounds indices
    loss = _MaskedPixelMSE.apply(pred, images.contiguous(), idx, normalize_target)
    loss._dig_mask_err = err      # device flag: some sample did not have exactly n masked patches (train_one_epoch reads it back)
    return loss


class _DevicePrefetch:
    """Iterates a loader one batch ahead: the (pinned) host tensors of batch n+1 are copied to the device on a side stream while step n
    computes, so the ~13 MB of crops per step cross PCIe under the previous step instead of in front of this one (the reference issues
    `.to(device, non_blocking=True)` on the compute stream, E:96-100).  Same batches, same order; anything that is not a tuple of CPU
    tensors passes through untouched."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, device
        self.stream = torch.cuda.Stream(device=device) if torch.cuda.is_available() else None

    def __len__(self):
        return len(self.loader)

    def _stage(self, item):
        batch = item[0]
        if (self.stream is None or not isinstance(batch, (tuple, list)) or len(batch) != 3
                or not all(isinstance(t, torch.Tensor) and not t.is_cuda for t in batch)):
            return item, None
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.stream):
            moved = tuple(t.to(self.device, non_blocking=True) for t in batch)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        for t in moved:
            t.record_stream(main)          # allocated on the copy stream, consumed on the compute stream
        return (moved,) + tuple(item[1:]), ev

    def __iter__(self):
        it = iter(self.loader)
        try:
            staged = self._stage(next(it))
        except StopIteration:
            return
        while staged is not None:
            cur, ev = staged
            try:
                staged = self._stage(next(it))
            except StopIteration:
                staged = None
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
            yield cur


def train_one_epoch(model, teacher_model, teacher_model_without_ddp, data_loader, word_data_loader, optimizer, device, epoch,
                    loss_scaler, max_norm=0, patch_size=16, normlize_target=True, log_writer=None, lr_scheduler=None,
                    start_steps=None, lr_schedule_values=None, wd_schedule_values=None, momentum_schedule=None, args=None):
    """Same contract as E:26-204.  `teacher_model`, `teacher_model_without_ddp`, `word_data_loader` and
    `momentum_schedule` are accepted and unused, exactly as in the reference."""
    if patch_size != 4:
        raise ops.DigError("patch_size must be 4 (pretrain_*_patch4_32x128), got %r" % (patch_size,))
    model.train()
    metric_logger = utils.MetricLogger(delimiter="  ")
    metric_logger.add_meter("lr", utils.SmoothedValue(window_size=1, fmt="{value:.6f}"))
    metric_logger.add_meter("min_lr", utils.SmoothedValue(window_size=1, fmt="{value:.6f}"))
    header = "Epoch: [{}]".format(epoch)
    print_freq = 100
    if args.num_view != 2:
        raise ops.DigError("num_view must be 2 (README.md:66)")
    n_it = len(data_loader)
    start_steps = 0 if start_steps is None else start_steps

    # contrast loss weight warm-up table, E:47-56
    if epoch == args.contrast_start_epoch:
        k = min(args.contrast_warmup_steps, n_it)
        w = np.linspace(0.0, args.loss_weight_contrast, k)
        if k < n_it:
            w = np.hstack([w, np.ones(n_it - k) * args.loss_weight_contrast])
    elif epoch > args.contrast_start_epoch:
        w = np.ones(n_it) * args.loss_weight_contrast
    else:
        w = np.zeros(n_it)

    pinned = [torch.empty(9, dtype=torch.float32).pin_memory() for _ in range(2)]
    pending = None

    def consume(rec):
        ev, slot, loss_scale_value, lr_max, lr_min, wd, = rec
        ev.synchronize()
        packed = slot.tolist()
        loss_value = packed[0]
        if not math.isfinite(loss_value):                                                             # E:148-150
            # The check runs one step late, but the non-finite step changed nothing: FusedAdamW's launch is guarded by the loss on the
            # device (dig_mt_adamw `guard`), so weights, moments and every checkpoint written before this point are clean.
            print("Loss is {}, stopping training".format(loss_value))
            sys.exit(1)
        if packed[8] != 0.0:
            raise ops.DigError("a batch did not mask the same number of patches in every sample (masking_generator.py:20)")
        metric_logger.update(loss_contrast=packed[1], q1_acc1=packed[3], q1_acc5=packed[4], q2_acc1=packed[5], q2_acc5=packed[6],
                             loss_pixel=packed[2], loss=loss_value, loss_scale=loss_scale_value)
        metric_logger.update(lr=lr_max, min_lr=lr_min, weight_decay=wd, grad_norm=packed[7])
        if log_writer is not None:
            log_writer.update(loss=loss_value, head="loss")
            log_writer.update(loss_scale=loss_scale_value, head="opt")
            log_writer.update(lr=lr_max, head="opt")
            log_writer.update(min_lr=lr_min, head="opt")
            log_writer.update(weight_decay=wd, head="opt")
            log_writer.update(grad_norm=packed[7], head="opt")
            log_writer.set_step()

    # measured: one GPU 22.46 -> 22.15 ms/step with the prefetch; two GPUs 23.03 ms without against 23.38 ms with it (the exchanges with the
    # peer already cover the copy there), so it defaults to on for a single process only (DIG_PREFETCH=0 / 1 overrides)
    prefetch = os.environ.get("DIG_PREFETCH", "1" if utils.get_world_size() == 1 else "0") != "0"
    feed = _DevicePrefetch(data_loader, device) if prefetch else data_loader
    for step, (batch, text, text_lens) in enumerate(metric_logger.log_every(feed, print_freq, header)):
        it = start_steps + step
        if lr_schedule_values is not None or wd_schedule_values is not None:                      # E:60-66
            for group in optimizer.param_groups:
                if lr_schedule_values is not None:
                    group["lr"] = lr_schedule_values[it] * group["lr_scale"]
                if wd_schedule_values is not None and group["weight_decay"] > 0:
                    group["weight_decay"] = wd_schedule_values[it]
        moco_m = utils.adjust_moco_momentum(epoch + 1.0 * step / n_it, args) if args.use_moco_m_cos else args.moco_m
        metric_logger.update(moco_m=moco_m)

        stage = getattr(args, "gpu_input_stage", None)
        if stage is not None and len(batch) == 2:
            # the loader ships the two views as uint8 [B,32,128,3]; ToTensor / Normalize / RandomGrayscale and the random
            # masks (datasets.py:27-49, dataset_image.py:39-52, masking_generator.py:12-46) run on the device (dig_b200/input_stage.py)
            world, rank = utils.get_world_size(), utils.get_rank()
            images, aug_images, mask = stage(batch[0], batch[1], sample0=(it * world + rank) * batch[0].shape[0], step=it)
        else:
            images, aug_images, mask = batch
            images = images.to(device, non_blocking=True)
            aug_images = aug_images.to(device, non_blocking=True)
            mask = mask.to(device, non_blocking=True).flatten(1).to(torch.bool).view(images.shape[0], args.num_view, -1)
        if args.only_mim_on_ori_img:
            mask[:, 1, :].fill_(0)                                                                    # E:103-104

        out = model(images, aug_images, mask, moco_m, args.only_mim_on_ori_img)
        contra = out["contra_loss"]
        # E:137-141: every masked view against the patches of the ORIGINAL image (E:83-111 builds all labels from `images`), 1/num_view each
        per_view = [masked_pixel_mse(v, images, mask[:, i], normalize_target=bool(normlize_target)) for i, v in enumerate(out["vis_out"])]
        loss_pixel = per_view[0] if len(per_view) == 1 else sum(per_view) * (1.0 / len(per_view))
        loss_pixel._dig_mask_err = sum(lp._dig_mask_err for lp in per_view)
        loss = contra * float(w[step]) + loss_pixel * float(args.loss_weight_pixel)

        optimizer.zero_grad()
        grad_norm = loss_scaler(loss, optimizer, clip_grad=max_norm, parameters=model.parameters(), create_graph=False)
        loss_scale_value = loss_scaler.state_dict()["scale"]

        # One packed device->host read per step (the reference blocks ~10 times, E:123-176), and it is consumed ONE STEP LATE: the copy
        # goes to a pinned buffer behind an event, and the meters / finite check of step n run after step n+1 has been enqueued, so the
        # host never drains the GPU queue inside the epoch.  Values, averages and the returned dict are unchanged; log lines lag a step.
        packed = torch.stack([loss.detach().float().reshape(()), contra.detach().float().reshape(()), loss_pixel.detach().reshape(()),
                              out["q1_acc1"].reshape(()), out["q1_acc5"].reshape(()), out["q2_acc1"].reshape(()),
                              out["q2_acc5"].reshape(()), grad_norm.to(loss.device).float().reshape(()),
                              loss_pixel._dig_mask_err.float().reshape(())])
        slot = pinned[step & 1]
        slot.copy_(packed, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        lrs = [g["lr"] for g in optimizer.param_groups]
        wds = [g["weight_decay"] for g in optimizer.param_groups if g["weight_decay"] > 0]
        this = (ev, slot, loss_scale_value, max(lrs), min(lrs), wds[-1] if wds else None)
        if pending is not None:
            consume(pending)
        pending = this
        if step == 0:          # the logger prints after the first iteration: give it real values (one synchronising read per epoch)
            consume(pending)
            pending = None
        if lr_scheduler is not None:
            lr_scheduler.step_update(start_steps + step)
        if step >= 1 and step % (args.eval_freq * 10) == 0 and getattr(args, "output_dir", None):
            from .checkpoint import save_model
            if pending is not None:        # never write a checkpoint of a step whose loss has not passed the finite check (E:148-150)
                consume(pending)
                pending = None
            save_model(args=args, model=model, model_without_ddp=getattr(model, "module", model), optimizer=optimizer,
                       loss_scaler=loss_scaler, epoch="{0}_{1}".format(epoch, step))
        sys.stdout.flush()

    if pending is not None:
        consume(pending)
    metric_logger.synchronize_between_processes()
    print("Averaged stats:", metric_logger)
    return {k: meter.global_avg for k, meter in metric_logger.meters.items()}
