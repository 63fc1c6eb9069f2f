#!/usr/bin/env python
"""bench.py -- DiG pre-training step throughput on B200 (metric of BASELINE.json: text-crops/sec, ViT-S/4 32x128).

    python bench.py --gpus N --steps K --warmup W            # dig_b200 arm (hand-written sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port of the reference, all host threads

One "step" = one full pre-training iteration on one batch of synthetic crops: two-view online forward +
momentum forward (with EMA update), InfoNCE + masked-pixel MSE, backward, gradient norm, AdamW.
`value` times K steps with inputs resident in HBM; `e2e` times K steps through the public API
(`dig_b200.engine.train_one_epoch`) with pinned HOST batches, H2D copies and the packed D2H metric read inside
the timed region.  Rank 0 prints ONE JSON line.
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
GFLOP_PER_CROP = {"pretrain_simmim_moco_ori_vit_small_patch4_32x128": 99.8, "pretrain_simmim_moco_ori_vit_base_patch4_32x128": 171.1,
                  "pretrain_simmim_moco_ori_vit_tiny_patch4_32x128": None}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dig", choices=["dig", "reference"])
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--batch", type=int, default=128, help="crops per GPU (BASELINE config 2/3)")
    ap.add_argument("--cpu-batch", type=int, default=8, help="crops per CPU-baseline step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
def cpu_reference_run(model_name, B, steps, warmup):
    """Times the oracle port of the reference step (fp32, forward + backward + grad-norm + AdamW) on the host cores."""
    import torch
    from oracle import restatement as R
    import dig_b200
    from dig_b200 import modeling  # noqa: F401  (parameter holders only: gives the reference's init and state-dict keys)
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = dig_b200.create_model(model_name, pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256,
                                  T=0.2, num_windows=4, encoder_type="vit", queue_size=65536, patchnet_name="no_patchtrans")
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


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    s = cpu_reference_run(a.model, a.cpu_batch, steps, warmup)
    v = a.cpu_batch / s
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "pretrain text-crops/sec", "value": v, "unit": "crops/s", "n_gpus": a.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "%s bs=%d/step on host CPU (bounded sample of the bs=128/GPU step), num_view=2, "
                                                        "mask_ratio=0.7, fwd+bwd+AdamW" % (a.model, a.cpu_batch)},
            "cpu_baseline": {"value": v, "unit": "crops/s", "cores": cores, "kind": "port",
                             "sample": "oracle/restatement.py OracleTrainer, B=%d, %d warm-up + %d timed steps, %d threads" % (
                                 a.cpu_batch, warmup, steps, cores)},
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


def gpu_arm(a):
    import torch
    import torch.distributed as dist
    import dig_b200
    from dig_b200 import modeling, ops  # noqa: F401
    from dig_b200.engine import masked_pixel_mse, train_one_epoch
    from dig_b200.optim import FusedAdamW
    from dig_b200.utils import NativeScalerWithGradNormCount

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # the reference's contrastive_loss needs an initialised group even on 1 GPU (M:449-453); ours does not, but DDP does
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ops.load()
    B = a.batch
    torch.manual_seed(0)
    model = dig_b200.create_model(a.model, pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2,
                                  num_windows=4, encoder_type="vit", queue_size=65536, patchnet_name="no_patchtrans")
    model.to(dev).train()
    net = model
    if world > 1:
        # (DIG_BENCH_NO_SYNCBN=1, experiment only: per-rank BatchNorm statistics, to size the cost of the SyncBN collectives)
        net = model if os.environ.get("DIG_BENCH_NO_SYNCBN") == "1" else torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
        if os.environ.get("DIG_BENCH_DDP", "dig") == "torch":
            # the reference runner's wrapper (R:391): works, but pays 2 x 183 per-parameter bucket copies and an unoverlapped all-reduce
            net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local],
                                                            find_unused_parameters=os.environ.get("DIG_BENCH_FIND_UNUSED", "1") != "0")
            if os.environ.get("DIG_BENCH_NOOP_ALLREDUCE") == "1":     # experiment only: how much of the step is DDP's gradient all-reduce?
                from torch.distributed.algorithms.ddp_comm_hooks.debugging_hooks import noop_hook
                net.register_comm_hook(None, noop_hook)
        else:
            from dig_b200.parallel import DigDataParallel
            net = DigDataParallel(net)     # flat-buffer gradient averaging overlapped with the backward (dig_b200/parallel.py)
    decay, no_decay = [], []
    for n, p in model.named_parameters():
        if p.requires_grad:
            (no_decay if (p.dim() == 1 or n.endswith(".bias")) else decay).append(p)
    lr = 1.5e-4 * B * world / 256
    opt = FusedAdamW([{"params": decay, "weight_decay": 0.05, "lr_scale": 1.0}, {"params": no_decay, "weight_decay": 0.0, "lr_scale": 1.0}],
                     lr=lr, betas=(0.9, 0.999), eps=1e-8)
    scaler = NativeScalerWithGradNormCount()
    args = make_args()
    img, aug, maskf = synthetic_batch(B, 1 + rank)
    img_d, aug_d = img.to(dev), aug.to(dev)
    mask_d = maskf.to(dev).flatten(1).to(torch.bool).view(B, 2, -1)
    mask_d[:, 1, :] = False

    def step_resident():
        out = net(img_d, aug_d, mask_d, 0.99, True)
        lp = masked_pixel_mse(out["vis_out"][0], img_d, mask_d[:, 0])
        loss = out["contra_loss"] * 0.1 + lp
        opt.zero_grad()
        scaler(loss, opt, clip_grad=None, parameters=model.parameters())
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(3, a.warmup)):
        last = step_resident()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        last = step_resident()
    e1.record()
    sync_all()
    launches = ops.launch_count() - n0
    ms = e0.elapsed_time(e1) / a.steps
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(last)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # ---- e2e through the public engine API with pinned host batches ----
    e2e = None
    if not a.no_e2e:
        host = [synthetic_batch(B, 100 + rank * 1000 + i, pin=True) for i in range(min(a.steps, 8))]
        loader = [([host[i % len(host)][0], host[i % len(host)][1], host[i % len(host)][2]], None, None) for i in range(a.steps)]
        warm = loader[:max(3, min(a.warmup, len(loader)))]
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            train_one_epoch(net, None, None, warm, None, opt, dev, 1, scaler, max_norm=None, patch_size=4, normlize_target=False,
                            start_steps=0, args=args)
            sync_all()
            e0.record()
            stats = train_one_epoch(net, None, None, loader, None, opt, dev, 1, scaler, max_norm=None, patch_size=4,
                                    normlize_target=False, start_steps=0, args=args)
            e1.record()
            sync_all()
        ems = e0.elapsed_time(e1) / a.steps
        t = torch.tensor([ems], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ems = float(t.item())
        h2d = img.numel() * 4 * 2 + maskf.numel() * 8
        e2e = {"value": B * world / (ems * 1e-3), "unit": "crops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * 4,
               "ms_per_step": ems, "loss_last_epoch_avg": stats.get("loss")}

    # ---- roofline of the dominant kernel family (tcgen05 GEMM), timed per launch with CUDA events in a separate pass ----
    roof = None
    if getattr(model, "_step", None) is not None:
        model._step._two_streams = False   # per-launch GEMM timing: one stream, so no co-running kernel is inside a timed interval
    ops.profile_gemm(True)      # every rank runs the two extra steps (they contain collectives); rank 0 reports
    for _ in range(2):
        step_resident()
    torch.cuda.synchronize()
    flops, gms, n, gbytes = ops.profile_gemm(False)
    sync_all()
    if rank == 0:
        peak, hbm, how = peaks()
        ach = flops / (gms * 1e-3) / 1e12 if gms > 0 else 0.0
        # DRAM traffic per launch of the same kernel family from the committed ncu pass (profiles/r1_step_metrics.json, written by
        # scripts/ncu_step_metrics.sh + summarize_step_metrics.py on a B200); algorithmic bytes are counted live from the shapes.
        traffic, tensor_pct, src, ncu_tf = None, None, None, None
        pj = os.path.join(ROOT, "profiles", "r1_step_metrics.json")
        if os.path.isfile(pj) and a.batch == 128 and a.model == MODEL:
            prof = json.load(open(pj))
            fam = prof.get("gemm_family", {})
            traffic, tensor_pct, src = fam.get("dram_bytes_per_launch"), fam.get("tensor_pipe_active_pct_time_weighted"), "profiles/r1_step_metrics.json"
            if fam.get("share_of_step") and prof.get("total_kernel_ms"):
                # the same FLOPs over the GEMM kernels' own durations in the committed ncu launch list (cold-cache, serialised): the
                # event-bracketed figure above also contains ~10 us of event/launch latency per launch
                ncu_tf = (flops / 2) / (fam["share_of_step"] * prof["total_kernel_ms"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05 / gemm2_bf16_tcgen05 (all encoder/head GEMM launches of the step)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": src, "algorithmic_bytes_per_launch": gbytes / max(n, 1),
                "hbm_frac_at_algorithmic_bytes": (gbytes / (gms * 1e-3) / 1e9) / hbm if gms > 0 else None,
                "ncu_tensor_pipe_active_pct": tensor_pct, "ncu_kernel_time_tflops": ncu_tf, "peak_source": how + " (sustained cuBLAS bf16)",
                "launches_timed": n, "avg_launch_ms": gms / max(n, 1), "gemm_share_of_step": (gms / 2) / ms}
        gf = GFLOP_PER_CROP.get(a.model)
        if gf:
            roof["whole_step_frac"] = (B / (ms * 1e-3)) * gf * 1e9 / (peak * 1e12)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        s = cpu_reference_run(a.model, a.cpu_batch, 2, 1)
        cores = os.cpu_count() or 1
        cpu = {"value": a.cpu_batch / s, "unit": "crops/s", "cores": cores, "kind": "port",
               "sample": "oracle/restatement.py OracleTrainer fp32, B=%d, 1 warm-up + 2 timed steps, %d threads" % (a.cpu_batch, cores)}

    if rank == 0:
        value = B * world / (ms * 1e-3)
        line = {"metric": "pretrain text-crops/sec", "value": value, "unit": "crops/s", "n_gpus": world, "steps": a.steps,
                "warmup": max(3, a.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "%s bs=%d/GPU num_view=2 mask_ratio=0.7 fwd+bwd+EMA+AdamW (BASELINE configs[%d])" % (
                    a.model, B, 1 if world == 1 else 2), "global_batch": B * world, "parallelism": "dp%d" % world, "dp_wrapper": (os.environ.get("DIG_BENCH_DDP", "dig") if world > 1 else None),
                    "l2": "per-step working set (>10 GB of saved activations) far exceeds the 126 MB L2; no explicit flush"},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "loss": loss_val}
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
