#!/usr/bin/env python
"""bench.py -- DiG pre-training step throughput on B200 (metric of BASELINE.json: text-crops/sec, ViT-S/4 32x128).

    python bench.py --gpus N --steps K --warmup W            # dig_b200 arm (hand-written sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the reference's own modules (staged baseline/_ref), all host threads

One "step" = one full pre-training iteration on one batch of synthetic crops: two-view online forward +
momentum forward (with EMA update), InfoNCE + masked-pixel MSE, backward, gradient norm, AdamW.
`value` times K steps with inputs resident in HBM; `e2e` times K steps through the public API
(`dig_b200.engine.train_one_epoch`) with pinned HOST batches, H2D copies and the packed D2H metric read inside
the timed region.  Rank 0 prints ONE JSON line.  Besides the contract keys it carries
  roofline            tcgen05 GEMM family (per-launch CUDA events) + `attention` sub-block (FLOP/s and MUFU-floor fractions)
  parity              one untimed step at THIS batch size against oracle/restatement.py in fp32 on the same GPU
  cpu_baseline        the reference step on the host cores (bounded sample)
  gpu_eager_baseline  the oracle port of the reference step, torch eager + bf16 autocast (cuBLASLt / ATen) on the same GPU
  vit_base            the ViT-B(512) bs=64 configuration (BASELINE configs[3]) measured the same way, fewer steps
  finetune            the fine-tuning step (BASELINE configs[4]): ViT-S encoder + tf_decoder, bs=256/GPU
  torch_ddp           (N > 1) the same step wrapped by torch DistributedDataParallel, what the unmodified runner constructs (R:391)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "pretrain_simmim_moco_ori_vit_small_patch4_32x128"
MODEL_BASE = "pretrain_simmim_moco_ori_vit_base_patch4_32x128"
GFLOP_PER_CROP = {MODEL: 99.8, MODEL_BASE: 171.1, "pretrain_simmim_moco_ori_vit_tiny_patch4_32x128": None}
KW = dict(pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2, num_windows=4, encoder_type="vit",
          queue_size=65536, patchnet_name="no_patchtrans")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dig", choices=["dig", "reference"])
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--batch", type=int, default=128, help="crops per GPU (BASELINE config 2/3)")
    ap.add_argument("--cpu-batch", type=int, default=16, help="crops per CPU-baseline step (bounded sample, BASELINE.md section 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip parity / gpu_eager_baseline / vit_base / torch_ddp")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1382.5), d.get("hbm_gbs", 6554.6), "measured"
    return 1400.0, 6650.0, "fallback"


def make_args(epochs=10):
    return types.SimpleNamespace(num_view=2, moco_m=0.99, use_moco_m_cos=1, epochs=epochs, contrast_start_epoch=0,
                                 contrast_warmup_steps=0, loss_weight_contrast=0.1, loss_weight_pixel=1.0, only_mim_on_ori_img=True,
                                 eval_freq=10 ** 9, output_dir=None)


def synthetic_batch(B, seed, pin=False):
    """images/aug ~ U(-1,1) fp32 [B,3,32,128]; mask float64 [B,2,256] with int(0.7*256)=179 ones per view."""
    import torch
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, 32, 128, generator=g) * 2 - 1
    aug = torch.rand(B, 3, 32, 128, generator=g) * 2 - 1
    mask = torch.zeros(B, 2, 256, dtype=torch.float64)
    perm = torch.rand(B, 2, 256, generator=g).argsort(dim=-1)[..., :179]
    mask.scatter_(2, perm, 1.0)
    if pin:
        img, aug, mask = img.pin_memory(), aug.pin_memory(), mask.pin_memory()
    return img, aug, mask


# ----------------------------------------------------------------------------------------------------- CPU arm
def _cpu_port_run(model_name, B, steps, warmup):
    """The oracle port of the reference step (fp32, forward + backward + grad-norm + AdamW) on the host cores."""
    import torch
    from oracle import restatement as R
    import dig_b200
    from dig_b200 import modeling  # noqa: F401  (parameter holders only: gives the reference's init and state-dict keys)
    torch.manual_seed(0)
    model = dig_b200.create_model(model_name, **KW)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    heads = model.encoder.num_heads
    del model
    tr = R.OracleTrainer(sd, heads, lr=1.5e-4 * B / 256, weight_decay=0.05)
    img, aug, mask = synthetic_batch(B, 1)
    mask = mask.bool()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        tr.step(img, aug, mask, 0.99)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def _cpu_reference_real_run(model_name, B, steps, warmup):
    """The UNMODIFIED reference (staged baseline/_ref or /root/reference, imported under oracle/ref_shims.py): its MoCo_ViT.forward
    (M:488-577), the engine's target / loss lines (E:83-144), autograd backward, utils.get_grad_norm_ (U:507-519) and
    custom_optim.AdamW (adamw.py:63-132) with optim_factory's decay / no-decay grouping -- fp32 on the host cores."""
    import torch
    import torch.distributed as dist
    from oracle import ref_shims
    from oracle import restatement as R
    M, E, U, AdamW = ref_shims.import_reference()
    created = False
    if not dist.is_initialized():          # contrastive_loss needs a group even at W = 1 (M:449-453)
        # an in-process store: under torchrun a tcp:// rendezvous is redirected to the elastic agent's store (TORCHELASTIC_USE_AGENT_STORE)
        # and a one-rank group of rank 0 alone would wait for it forever
        dist.init_process_group("gloo", store=dist.HashStore(), rank=0, world_size=1)
        created = True
    try:
        model = ref_shims.create_reference_model(model_name, seed=0)
        decay = [p for n, p in model.named_parameters() if p.requires_grad and not (p.dim() == 1 or n.endswith(".bias"))]
        no_decay = [p for n, p in model.named_parameters() if p.requires_grad and (p.dim() == 1 or n.endswith(".bias"))]
        opt = AdamW([{"params": decay, "weight_decay": 0.05}, {"params": no_decay, "weight_decay": 0.0}], lr=1.5e-4 * B / 256,
                    betas=(0.9, 0.999), eps=1e-8)
        img, aug, maskf = synthetic_batch(B, 1)
        mk = maskf.bool()
        mk[:, 1, :] = False
        times = []
        with ref_shims.cpu_patches():
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                labels = R.build_targets(img, mk)[0]                         # E:83-111 (restated: the engine inlines these lines)
                out = model(img, aug, mk, 0.99, True)
                loss = out["contra_loss"] * 0.1 + torch.nn.functional.mse_loss(out["vis_out"][0], labels)     # E:120-144
                opt.zero_grad()
                loss.backward()
                U.get_grad_norm_(model.parameters())
                opt.step()
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
        return sum(times) / len(times)
    finally:
        if created:
            dist.destroy_process_group()


def cpu_reference_run(model_name, B, steps, warmup):
    """-> (seconds per step, kind, sample description)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    note = ""
    try:
        from oracle import ref_shims
        if ref_shims.reference_available():
            s = _cpu_reference_real_run(model_name, B, steps, warmup)
            return s, "reference", ("unmodified reference modules (MoCo_ViT.forward + autograd + get_grad_norm_ + custom_optim.AdamW) imported "
                                    "from %s, fp32, B=%d, %d warm-up + %d timed steps, %d threads" % (
                                        os.path.relpath(ref_shims.REFERENCE_ROOT, ROOT) if ref_shims.REFERENCE_ROOT.startswith(ROOT)
                                        else ref_shims.REFERENCE_ROOT, B, warmup, steps, cores))
        note = " (reference not staged: baseline/_ref absent)"
    except Exception as e:      # fall back to the port rather than lose the baseline
        note = " (real reference failed: %r)" % (e,)
    s = _cpu_port_run(model_name, B, steps, warmup)
    return s, "port", "oracle/restatement.py OracleTrainer fp32, B=%d, %d warm-up + %d timed steps, %d threads%s" % (B, warmup, steps, cores, note)


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    s, kind, sample = cpu_reference_run(a.model, a.cpu_batch, steps, warmup)
    v = a.cpu_batch / s
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "pretrain text-crops/sec", "value": v, "unit": "crops/s", "n_gpus": a.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "%s bs=%d/step on host CPU (bounded sample of the bs=128/GPU step), num_view=2, "
                                                        "mask_ratio=0.7, fwd+bwd+AdamW" % (a.model, a.cpu_batch)},
            "cpu_baseline": {"value": v, "unit": "crops/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


class Workload:
    """Model + optimizer + resident synthetic batch for one configuration; `step()` is one full pre-training iteration."""

    def __init__(self, model_name, B, dev, rank, world, wrapper):
        import torch
        import dig_b200
        from dig_b200 import modeling  # noqa: F401
        from dig_b200.optim import FusedAdamW
        from dig_b200.utils import NativeScalerWithGradNormCount
        self.B, self.dev, self.world = B, dev, world
        torch.manual_seed(0)
        model = dig_b200.create_model(model_name, **KW)
        model.to(dev).train()
        net = model
        if world > 1:
            # (DIG_BENCH_NO_SYNCBN=1, experiment only: per-rank BatchNorm statistics, to size the cost of the SyncBN exchanges)
            net = model if os.environ.get("DIG_BENCH_NO_SYNCBN") == "1" else torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
            if wrapper == "torch":
                # the reference runner's wrapper (R:391)
                net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[dev.index], find_unused_parameters=True)
            elif wrapper == "none":
                pass    # experiment only: no gradient averaging (with DIG_BENCH_NO_SYNCBN=1 the ranks are nearly independent replicas)
            else:
                from dig_b200.parallel import DigDataParallel
                net = DigDataParallel(net)     # flat-buffer gradient averaging overlapped with the backward (dig_b200/parallel.py)
        decay, no_decay = [], []
        for n, p in model.named_parameters():
            if p.requires_grad:
                (no_decay if (p.dim() == 1 or n.endswith(".bias")) else decay).append(p)
        self.opt = FusedAdamW([{"params": decay, "weight_decay": 0.05, "lr_scale": 1.0}, {"params": no_decay, "weight_decay": 0.0, "lr_scale": 1.0}],
                              lr=1.5e-4 * B * world / 256, betas=(0.9, 0.999), eps=1e-8)
        self.scaler = NativeScalerWithGradNormCount()
        self.model, self.net = model, net
        self.img, self.aug, self.maskf = synthetic_batch(B, 1 + rank)
        self.img_d, self.aug_d = self.img.to(dev), self.aug.to(dev)
        self.mask_d = self.maskf.to(dev).flatten(1).to(torch.bool).view(B, 2, -1)
        self.mask_d[:, 1, :] = False

    def step(self):
        from dig_b200.engine import masked_pixel_mse
        out = self.net(self.img_d, self.aug_d, self.mask_d, 0.99, True)
        lp = masked_pixel_mse(out["vis_out"][0], self.img_d, self.mask_d[:, 0])
        loss = out["contra_loss"] * 0.1 + lp
        self.opt.zero_grad()
        self.scaler(loss, self.opt, clip_grad=None, parameters=self.model.parameters())
        return loss


class FinetuneWorkload:
    """BASELINE configs[4]: fine-tune simmim_vit_small_patch4_32x128 + tf_decoder, bs=256/GPU (dropout 0): DigRecModel forward,
    SeqCrossEntropyLoss, backward, grad-norm, fused AdamW on a resident synthetic batch (images U(-1,1), 25-position labels)."""

    def __init__(self, B, dev, rank, world):
        import torch
        from dig_b200.finetune import DigRecModel
        from dig_b200.optim import FusedAdamW
        from dig_b200.utils import NativeScalerWithGradNormCount
        self.B, self.dev = B, dev
        torch.manual_seed(0)
        model = DigRecModel("simmim_vit_small_patch4_32x128").to(dev).train()
        self.model = self.net = model
        if world > 1:      # run_class_finetuning.py:497 (mask_token gets no gradient here)
            self.net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], find_unused_parameters=True)
        decay, no_decay = [], []
        for n, p in model.named_parameters():
            if p.requires_grad:
                (no_decay if (p.dim() == 1 or n.endswith(".bias")) else decay).append(p)
        self.opt = FusedAdamW([{"params": decay, "weight_decay": 0.05, "lr_scale": 1.0}, {"params": no_decay, "weight_decay": 0.0, "lr_scale": 1.0}],
                              lr=1e-4 * B * world / 256, betas=(0.9, 0.999), eps=1e-8)
        self.scaler = NativeScalerWithGradNormCount()
        g = torch.Generator().manual_seed(7 + rank)
        self.img = (torch.rand(B, 3, 32, 128, generator=g) * 2 - 1).to(dev)
        lens = torch.randint(1, 26, (B,), generator=g)
        tgt = torch.randint(0, 94, (B, 25), generator=g)
        pos = torch.arange(25)[None, :]
        tgt = torch.where(pos == (lens[:, None] - 1), torch.full_like(tgt, 94), tgt)
        tgt = torch.where(pos >= lens[:, None], torch.full_like(tgt, 95), tgt)
        self.tgt, self.lens = tgt.to(dev), lens.to(dev)

    def step(self):
        from dig_b200.finetune import seq_cross_entropy
        logits = self.net((self.img, self.tgt, self.lens))[0]
        loss, _ = seq_cross_entropy(logits, self.tgt, self.lens)
        self.opt.zero_grad()
        self.scaler(loss, self.opt, clip_grad=None, parameters=self.model.parameters())
        return loss


def sync_all(world):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def timed_steps(wl, steps, warmup, world):
    """W warm-up steps, then K steps bracketed by barrier + synchronize, CUDA events, max over ranks -> (ms/step, last loss, launches)."""
    import torch
    import torch.distributed as dist
    from dig_b200 import ops
    last = None
    for _ in range(max(3, warmup)):
        last = wl.step()
    sync_all(world)
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        last = wl.step()
    e1.record()
    sync_all(world)
    launches = ops.launch_count() - n0
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=wl.dev)
    if world > 1:
        if os.environ.get("DIG_BENCH_RANK_TIMES") == "1":      # experiment: every rank's own figure (rank skew)
            all_t = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(all_t, t)
            if dist.get_rank() == 0:
                print("per-rank ms/step:", ["%.3f" % float(x.item()) for x in all_t], file=sys.stderr)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), float(last), launches


def parity_block(model_name, B, dev):
    """One untimed step of a fresh model (same seed) at the bench batch size against oracle/restatement.py run in fp32 on this GPU:
    relative loss differences and the worst per-tensor gradient cosine (encoder / heads)."""
    import torch
    import dig_b200
    from dig_b200 import modeling  # noqa: F401
    from dig_b200.engine import masked_pixel_mse
    from oracle import restatement as R
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    model = dig_b200.create_model(model_name, **KW).train()
    sd = {k: v.detach().clone().to(dev) for k, v in model.state_dict().items()}
    names = R.trainable_names(sd)
    for n in names:
        sd[n].requires_grad_(True)
    img, aug, mask = R.synthetic_batch(B, seed=1)
    img, aug, mask = img.to(dev), aug.to(dev), mask.to(dev)
    loss_o, out_o, lpix_o = R.step_losses(sd, img, aug, mask, 0.99, model.encoder.num_heads)
    grads_o = dict(zip(names, torch.autograd.grad(loss_o, [sd[n] for n in names], allow_unused=True)))
    ref = (float(loss_o), float(out_o["contra_loss"]), float(lpix_o))
    del out_o, loss_o, lpix_o, sd
    torch.cuda.empty_cache()
    mk = mask.clone()
    mk[:, 1, :] = False
    model.to(dev)
    out = model(img, aug, mk, 0.99, True)
    lpix = masked_pixel_mse(out["vis_out"][0], img, mk[:, 0])
    loss = out["contra_loss"] * 0.1 + lpix
    loss.backward()
    torch.cuda.synchronize()
    got = (float(loss), float(out["contra_loss"]), float(lpix))
    cos = {}
    for n, p in model.named_parameters():
        go = grads_o.get(n)
        if go is not None and p.grad is not None and float(go.norm()) > 0:
            cos[n] = float(torch.nn.functional.cosine_similarity(p.grad.float().flatten(), go.flatten(), dim=0))
    enc = [c for n, c in cos.items() if n.startswith("encoder.")]
    hd = [c for n, c in cos.items() if not n.startswith("encoder.")]
    rel = [abs(g - r) / abs(r) for g, r in zip(got, ref)]
    return {"against": "oracle/restatement.py fp32 on the same GPU (TF32 off), same seed / inputs, batch %d" % B, "loss_rel": rel[0],
            "contra_loss_rel": rel[1], "pixel_loss_rel": rel[2], "loss": got[0], "loss_oracle": ref[0], "grad_cos_min": min(enc + hd),
            "grad_cos_min_encoder": min(enc), "grad_cos_min_heads": min(hd), "grad_tensors": len(cos), "tolerance": "losses 1e-3 relative"}


def gpu_eager_baseline(model_name, B, dev, steps=3):
    """The oracle port of the reference step (the reference's module structure as plain PyTorch ops) run eagerly by torch on this GPU under
    bf16 autocast: cuBLASLt + ATen sm_100 kernels, forward + backward + grad-norm + AdamW on the same synthetic batch."""
    import torch
    import dig_b200
    from dig_b200 import modeling  # noqa: F401
    from oracle import restatement as R
    torch.manual_seed(0)
    model = dig_b200.create_model(model_name, **KW)
    sd = {k: v.detach().clone().to(dev) for k, v in model.state_dict().items()}
    heads = model.encoder.num_heads
    del model
    tr = R.OracleTrainer(sd, heads, lr=1.5e-4 * B / 256, weight_decay=0.05)
    img, aug, mask = R.synthetic_batch(B, seed=1)
    img, aug, mask = img.to(dev), aug.to(dev), mask.to(dev)

    def one():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return tr.step(img, aug, mask, 0.99)
    for _ in range(2):
        one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    del tr, sd
    torch.cuda.empty_cache()
    return {"value": B / dt, "unit": "crops/s", "ms_per_step": dt * 1e3, "kind": "port",
            "what": "oracle/restatement.py (the reference's module structure as plain torch ops) eager on this GPU, torch.autocast bf16, "
                    "fwd+bwd+grad-norm+AdamW, B=%d, 2 warm-up + %d timed steps" % (B, steps)}


def roofline_block(wl, a, ms, rank, world):
    """GEMM family + attention, timed per launch with CUDA events in two extra single-stream steps (every rank runs them: they contain
    collectives; rank 0 reports)."""
    import torch
    from dig_b200 import ops
    if getattr(wl.model, "_step", None) is not None:
        wl.model._step._two_streams = False   # per-launch timing: one stream, so no co-running kernel is inside a timed interval
    ops.profile_gemm(True)
    ops.profile_attention(True)
    for _ in range(2):
        wl.step()
    torch.cuda.synchronize()
    flops, gms, n, gbytes = ops.profile_gemm(False)
    att = ops.profile_attention(False)
    if getattr(wl.model, "_step", None) is not None:
        wl.model._step._two_streams = os.environ.get("DIG_TWO_STREAMS", "1") != "0"
    sync_all(world)
    if rank != 0:
        return None
    peak, hbm, how = peaks()
    ach = flops / (gms * 1e-3) / 1e12 if gms > 0 else 0.0
    # DRAM traffic per launch of the same kernel family from the committed ncu pass (scripts/ncu_step_metrics.sh +
    # summarize_step_metrics.py on a B200); algorithmic bytes are counted live from the shapes.
    traffic, tensor_pct, src, ncu_tf = None, None, None, None
    for tag in ("r2", "r1"):
        pj = os.path.join(ROOT, "profiles", "%s_step_metrics.json" % tag)
        if os.path.isfile(pj) and a.batch == 128 and a.model == MODEL:
            prof = json.load(open(pj))
            fam = prof.get("gemm_family", {})
            traffic, tensor_pct, src = fam.get("dram_bytes_per_launch"), fam.get("tensor_pipe_active_pct_time_weighted"), "profiles/%s_step_metrics.json" % tag
            if fam.get("share_of_step") and prof.get("total_kernel_ms"):
                # the same FLOPs over the GEMM kernels' own durations in the committed ncu launch list (cold-cache, serialised): the
                # event-bracketed figure above also contains ~10 us of event/launch latency per launch
                ncu_tf = (flops / 2) / (fam["share_of_step"] * prof["total_kernel_ms"] * 1e-3) / 1e12
            break
    roof = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05 / gemm2_bf16_tcgen05 (all encoder/head GEMM launches of the step)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
            "traffic_source": src, "algorithmic_bytes_per_launch": gbytes / max(n, 1),
            "hbm_frac_at_algorithmic_bytes": (gbytes / (gms * 1e-3) / 1e9) / hbm if gms > 0 else None,
            "ncu_tensor_pipe_active_pct": tensor_pct, "ncu_kernel_time_tflops": ncu_tf, "peak_source": how + " (sustained cuBLAS bf16)",
            "launches_timed": n, "avg_launch_ms": gms / max(n, 1), "gemm_share_of_step": (gms / 2) / ms}
    gf = GFLOP_PER_CROP.get(a.model)
    if gf:
        roof["whole_step_frac"] = (a.batch / (ms * 1e-3)) * gf * 1e9 / (peak * 1e12)
    # attention: achieved fraction of the attention-GEMM roofline (north_star) and of the exp2 (MUFU) floor: 16 ex2 / clk / SM
    sm_max = 1965.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        sm_max = json.load(open(pk)).get("sm_max_mhz", sm_max)
    mufu_per_s = 16.0 * 148 * sm_max * 1e6
    ab = {}
    for kind in ("fwd", "bwd"):
        fl, tms, cnt, exps = att[kind]
        if cnt:
            tf = fl / (tms * 1e-3) / 1e12
            ab[kind] = {"achieved": tf, "unit": "TFLOP/s", "frac": tf / peak, "launches_timed": cnt, "avg_launch_ms": tms / cnt,
                        "mufu_floor_ms": exps / cnt / mufu_per_s * 1e3, "mufu_floor_frac": (exps / cnt / mufu_per_s * 1e3) / (tms / cnt),
                        "algorithmic_flop_per_launch": fl / cnt}
    ab["peak"] = peak
    ab["note"] = ("algorithmic FLOPs 4 (fwd) / 10 (bwd) x 256 x 256 x 64 per (sequence, head), no recompute counted; mufu_floor = one exp2 "
                  "per score at 16/clk/SM x 148 SMs x %.0f MHz" % sm_max)
    roof["attention"] = ab
    return roof


def gpu_arm(a):
    import torch
    import torch.distributed as dist
    from dig_b200 import ops
    from dig_b200.engine import train_one_epoch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ops.load()
    B = a.batch
    wrapper = os.environ.get("DIG_BENCH_DDP", "dig")
    wl = Workload(a.model, B, dev, rank, world, wrapper)

    sampler = ClockSampler(local)
    for _ in range(max(3, a.warmup)):
        wl.step()
    sync_all(world)
    if rank == 0:
        sampler.start()
    ms, loss_val, launches = timed_steps(wl, a.steps, 0, world)
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e through the public engine API with pinned host batches ----
    e2e = None
    if not a.no_e2e:
        args = make_args()
        host = [synthetic_batch(B, 100 + rank * 1000 + i, pin=True) for i in range(min(a.steps, 8))]
        loader = [([host[i % len(host)][0], host[i % len(host)][1], host[i % len(host)][2]], None, None) for i in range(a.steps)]
        warm = loader[:max(3, min(a.warmup, len(loader)))]
        import contextlib, io
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with contextlib.redirect_stdout(io.StringIO()):
            train_one_epoch(wl.net, None, None, warm, None, wl.opt, dev, 1, wl.scaler, max_norm=None, patch_size=4, normlize_target=False,
                            start_steps=0, args=args)
            sync_all(world)
            e0.record()
            stats = train_one_epoch(wl.net, None, None, loader, None, wl.opt, dev, 1, wl.scaler, max_norm=None, patch_size=4,
                                    normlize_target=False, start_steps=0, args=args)
            e1.record()
            sync_all(world)
        t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ems = float(t.item())
        h2d = wl.img.numel() * 4 * 2 + wl.maskf.numel() * 8
        e2e = {"value": B * world / (ems * 1e-3), "unit": "crops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 9 * 4,
               "ms_per_step": ems, "loss_last_epoch_avg": stats.get("loss")}

    roof = roofline_block(wl, a, ms, rank, world)

    extras = {}
    if not a.no_extras:
        # the same step under torch DistributedDataParallel: what the unmodified runner constructs (R:390-391)
        if world > 1 and wrapper != "torch" and os.environ.get("DIG_BENCH_TORCH_DDP", "1") != "0":
            try:
                wl2 = Workload(a.model, B, dev, rank, world, "torch")
                ms2, _, _ = timed_steps(wl2, max(5, a.steps // 2), 3, world)
                extras["torch_ddp"] = {"value": B * world / (ms2 * 1e-3), "unit": "crops/s", "ms_per_step": ms2,
                                       "what": "same step, SyncBatchNorm + torch DistributedDataParallel(find_unused_parameters=True) "
                                               "instead of dig_b200.parallel.DigDataParallel"}
                del wl2
            except Exception as e:
                extras["torch_ddp"] = {"error": repr(e)}
            sync_all(world)
        if rank == 0 and world == 1:
            for key, fn in (("parity", lambda: parity_block(a.model, B, dev)),
                            ("gpu_eager_baseline", lambda: gpu_eager_baseline(a.model, B, dev))):
                try:
                    extras[key] = fn()
                except Exception as e:
                    extras[key] = {"error": repr(e)}
                torch.cuda.empty_cache()
        # BASELINE configs[3]: ViT-B(512) bs=64/GPU, same measurement with fewer steps (every rank takes part)
        if a.model == MODEL and os.environ.get("DIG_BENCH_VIT_BASE", "1") != "0":
            try:
                wlb = Workload(MODEL_BASE, 64, dev, rank, world, wrapper)
                msb, lossb, _ = timed_steps(wlb, max(5, a.steps // 2), 3, world)
                peak = peaks()[0]
                extras["vit_base"] = {"value": 64 * world / (msb * 1e-3), "unit": "crops/s", "ms_per_step": msb, "loss": lossb,
                                      "config": {"workload": "%s bs=64/GPU (BASELINE configs[3]), resident inputs" % MODEL_BASE,
                                                 "global_batch": 64 * world},
                                      "whole_step_frac": (64 / (msb * 1e-3)) * GFLOP_PER_CROP[MODEL_BASE] * 1e9 / (peak * 1e12)}
                del wlb
            except Exception as e:
                extras["vit_base"] = {"error": repr(e)}
            sync_all(world)

    # BASELINE configs[4]: the fine-tuning step, bs=256/GPU, every rank takes part
    if not a.no_extras and a.model == MODEL and os.environ.get("DIG_BENCH_FINETUNE", "1") != "0":
        try:
            wlf = FinetuneWorkload(256, dev, rank, world)
            msf, lossf, _ = timed_steps(wlf, max(5, a.steps // 2), 3, world)
            gf = 43.1     # GFLOP per sample: encoder 3 x 12.089 + linear_norm / decoder / classifier 3 x 2.27 (2 FLOP per MAC, bwd = 2 x fwd)
            extras["finetune"] = {"value": 256 * world / (msf * 1e-3), "unit": "samples/s", "ms_per_step": msf, "loss": lossf,
                                  "config": {"workload": "simmim_vit_small_patch4_32x128 + tf_decoder bs=256/GPU, max_len 25, dropout 0, "
                                                         "fwd+bwd+AdamW (BASELINE configs[4]), resident inputs", "global_batch": 256 * world,
                                             "dp_wrapper": "torch DistributedDataParallel" if world > 1 else None},
                                  "whole_step_frac": (256 / (msf * 1e-3)) * gf * 1e9 / (peaks()[0] * 1e12)}
            del wlf
        except Exception as e:
            extras["finetune"] = {"error": repr(e)}
        sync_all(world)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            s, kind, sample = cpu_reference_run(a.model, a.cpu_batch, 2, 1)
            cpu = {"value": a.cpu_batch / s, "unit": "crops/s", "cores": os.cpu_count() or 1, "kind": kind, "sample": sample}
        except Exception as e:
            cpu = {"error": repr(e)}

    if rank == 0:
        value = B * world / (ms * 1e-3)
        cfg_idx = 3 if a.model == MODEL_BASE else (1 if world == 1 else 2)
        line = {"metric": "pretrain text-crops/sec", "value": value, "unit": "crops/s", "n_gpus": world, "steps": a.steps,
                "warmup": max(3, a.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "%s bs=%d/GPU num_view=2 mask_ratio=0.7 fwd+bwd+EMA+AdamW (BASELINE configs[%d])" % (
                    a.model, B, cfg_idx), "global_batch": B * world, "parallelism": "dp%d" % world, "dp_wrapper": (wrapper if world > 1 else None),
                    "l2": "per-step working set (>10 GB of saved activations) far exceeds the 126 MB L2; no explicit flush"},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "loss": loss_val}
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        gpu_arm(a)


if __name__ == "__main__":
    main()
