"""Drop-in for `engine_for_finetuning.train_one_epoch` (reference engine_for_finetuning.py:54-210) on the fused fine-tuning step
(dig_b200/finetune.py).  Same signature and meter set; built for the AMP-scaler path the README runs
(`loss_scaler` given, no DeepSpeed, no mixup, no teacher).  As in the pre-training engine the per-step `.item()` reads (E-ft:107-176)
are one packed device->host copy consumed one step late, and a non-finite loss never reaches the weights (guarded optimizer launch).
`class_acc` is the word accuracy the reference computes through evaluation_metric (every position up to and including <EOS> correct),
evaluated on the label ids on the device."""
import math
import sys

import torch

from . import ops, utils
from .finetune import seq_cross_entropy


def train_one_epoch(model, criterion, data_loader, optimizer, device, epoch, loss_scaler, max_norm=0, model_ema=None, mixup_fn=None,
                    log_writer=None, start_steps=None, lr_schedule_values=None, wd_schedule_values=None, num_training_steps_per_epoch=None,
                    update_freq=None, args=None, data_loader_val=None, max_accuracy=0., criterion_aux=None, teacher_model=None):
    if loss_scaler is None or mixup_fn is not None or model_ema is not None or teacher_model is not None:
        raise NotImplementedError("the fused fine-tuning engine is built for the loss_scaler path without mixup / EMA / teacher (README.md:91-112)")
    update_freq = update_freq or 1
    if update_freq != 1:
        raise NotImplementedError("gradient accumulation (update_freq > 1) is not built")
    model.train(True)
    metric_logger = utils.MetricLogger(delimiter="  ")
    metric_logger.add_meter("lr", utils.SmoothedValue(window_size=1, fmt="{value:.6f}"))
    metric_logger.add_meter("min_lr", utils.SmoothedValue(window_size=1, fmt="{value:.6f}"))
    header = "Epoch: [{}]".format(epoch)
    start_steps = 0 if start_steps is None else start_steps
    n_steps = num_training_steps_per_epoch if num_training_steps_per_epoch is not None else len(data_loader)
    optimizer.zero_grad()
    pinned = [torch.empty(3, dtype=torch.float32).pin_memory() for _ in range(2)]
    pending = None

    def consume(rec):
        ev, slot, scale, lr_max, lr_min, wd = rec
        ev.synchronize()
        loss_value, acc, gn = slot.tolist()
        if not math.isfinite(loss_value):                                                           # E-ft:107-111
            print("Loss is {}, stopping training".format(loss_value))
            sys.exit(1)
        metric_logger.update(loss=loss_value, class_acc=acc, loss_scale=scale)
        metric_logger.update(lr=lr_max, min_lr=lr_min, weight_decay=wd, grad_norm=gn)
        if log_writer is not None:
            log_writer.update(loss=loss_value, head="loss")
            log_writer.update(class_acc=acc, head="loss")
            log_writer.update(loss_scale=scale, head="opt")
            log_writer.update(lr=lr_max, head="opt")
            log_writer.update(min_lr=lr_min, head="opt")
            log_writer.update(weight_decay=wd, head="opt")
            log_writer.update(grad_norm=gn, head="opt")
            log_writer.set_step()

    for step, data in enumerate(metric_logger.log_every(data_loader, 100, header)):
        samples, targets, tgt_lens = data[:3]
        if step >= n_steps:
            continue
        it = start_steps + step
        if lr_schedule_values is not None or wd_schedule_values is not None:                       # E-ft:88-93
            for group in optimizer.param_groups:
                if lr_schedule_values is not None:
                    group["lr"] = lr_schedule_values[it] * group.get("lr_scale", 1.0)
                if wd_schedule_values is not None and group["weight_decay"] > 0:
                    group["weight_decay"] = wd_schedule_values[it]
        samples = samples.to(device, non_blocking=True)
        targets = targets.to(device, non_blocking=True)
        tgt_lens = tgt_lens.to(device, non_blocking=True)
        outputs = model((samples, targets, tgt_lens))[0]                                            # train_class_batch, E-ft:26-47
        loss, pred = seq_cross_entropy(outputs, targets, tgt_lens)
        grad_norm = loss_scaler(loss, optimizer, clip_grad=max_norm, parameters=model.parameters(), create_graph=False, update_grad=True)
        optimizer.zero_grad()
        scale = loss_scaler.state_dict()["scale"]
        T = targets.shape[1]
        valid = torch.arange(T, device=targets.device)[None, :] < tgt_lens.reshape(-1, 1)
        acc = ((pred == targets) | ~valid).all(dim=1).float().mean()
        packed = torch.stack([loss.detach().float().reshape(()), acc, grad_norm.to(loss.device).float().reshape(())])
        slot = pinned[step & 1]
        slot.copy_(packed, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        lrs = [g["lr"] for g in optimizer.param_groups]
        wds = [g["weight_decay"] for g in optimizer.param_groups if g["weight_decay"] > 0]
        this = (ev, slot, scale, max(lrs), min(lrs), wds[-1] if wds else None)
        if pending is not None:
            consume(pending)
        pending = this
        if step == 0:
            consume(pending)
            pending = None
        sys.stdout.flush()
    if pending is not None:
        consume(pending)
    metric_logger.synchronize_between_processes()
    print("Averaged stats:", metric_logger)
    stats = {k: meter.global_avg for k, meter in metric_logger.meters.items()}
    stats.update({"max_accuracy": max_accuracy})
    return stats


# ---- evaluation (reference engine_for_finetuning.py:212-285; metrics of evaluation_metric/metrics.py:14-96 restated on label ids) ----
def _strings(ids, dataset):
    """evaluation_metric/metrics.py:19-64 get_str_list for one id matrix: characters up to (not including) <EOS>, <UNKNOWN> dropped,
    then only [0-9a-zA-Z] kept and lower-cased (_normalize_text)."""
    import string
    keep = set(string.digits + string.ascii_letters)
    eos, unk = dataset.class_to_idx["EOS"], dataset.class_to_idx["UNKNOWN"]
    out = []
    for row in ids.tolist():
        chars = []
        for v in row:
            if v == eos:
                break
            if v != unk:
                chars.append(dataset.idx_to_class[v])
        out.append("".join(c for c in "".join(chars) if c in keep).lower())
    return out


def word_accuracy(pred_ids, target, dataset):
    """evaluation_metric.Accuracy (metrics.py:76-81): fraction of samples whose normalised strings are equal."""
    p, t = _strings(pred_ids, dataset), _strings(target, dataset)
    return sum(a == b for a, b in zip(p, t)) / max(len(p), 1)


def recognition_fmeasure(pred_ids, target, dataset):
    """evaluation_metric.recognition_f_measure (metrics.py:83-100): mean F-measure over the SETS of characters of prediction and target."""
    fs = []
    for pred, targ in zip(_strings(pred_ids, dataset), _strings(target, dataset)):
        pc, tc = set(pred), set(targ)
        right = float(len(pc & tc))
        p, r = right / (len(pc) + 1e-5), right / (len(tc) + 1e-5)
        fs.append(2 * p * r / (p + r + 1e-5))
    return sum(fs) / max(len(fs), 1)


@torch.no_grad()
def evaluate(data_loader, model, device, args=None):
    """Same contract as the reference's evaluate (engine_for_finetuning.py:212-285) for the tf_decoder model: eval mode (greedy decoding on
    the fused kernels, dig_b200/finetune.py), the criterion applied to the decoder's step PROBABILITIES exactly as the reference does
    (E-ft:240), word accuracy and the character-set F-measure; returns the meters' global averages."""
    from .finetune import SeqCrossEntropyLoss
    if getattr(args, "beam_width", 0):
        raise NotImplementedError("beam search (models/decoder.py:252-330) is not built; evaluate with --beam_width 0")
    criterion = SeqCrossEntropyLoss()
    metric_logger = utils.MetricLogger(delimiter="  ")
    header = "Test:"
    model.eval()
    for batch in metric_logger.log_every(data_loader, 10, header):
        images, target, lens = batch[0], batch[1], batch[-1]
        images = images.to(device, non_blocking=True)
        target = target.to(device, non_blocking=True)
        lens = lens.to(device, non_blocking=True)
        output = model((images, target, lens))[0]
        loss = criterion(output, target, lens)
        pred_ids = output.argmax(-1)
        dataset = data_loader.dataset
        bs = images.shape[0]
        metric_logger.update(loss=float(loss))
        metric_logger.meters["acc"].update(word_accuracy(pred_ids.cpu(), target.cpu(), dataset), n=bs)
        metric_logger.meters["recognition_fmeasure"].update(recognition_fmeasure(pred_ids.cpu(), target.cpu(), dataset), n=bs)
    metric_logger.synchronize_between_processes()
    print("* {} images, Acc {acc.global_avg:.4f} loss {losses.global_avg:.4f} Rec_fmeasure {rec_f.global_avg:.4f}".format(
        metric_logger.meters["acc"].count, acc=metric_logger.meters["acc"], losses=metric_logger.meters["loss"],
        rec_f=metric_logger.meters["recognition_fmeasure"]))
    return {k: meter.global_avg for k, meter in metric_logger.meters.items()}
