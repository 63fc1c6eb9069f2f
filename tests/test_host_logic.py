"""Host-side logic of the engine (no GPU): schedules, meters, loss-weight warm-up, rank layout under a 2-process gloo group."""
import math
import os
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dig_b200 import utils


def test_cosine_scheduler_matches_reference_formula():
    s = utils.cosine_scheduler(1.0, 0.1, epochs=3, niter_per_ep=10, warmup_epochs=1, start_warmup_value=0.0)
    assert len(s) == 30 and s[0] == 0.0 and s[9] == pytest.approx(1.0)
    it = np.arange(20)
    assert np.allclose(s[10:], 0.1 + 0.45 * (1 + np.cos(np.pi * it / 20)))
    s2 = utils.cosine_scheduler(1.0, 0.1, epochs=3, niter_per_ep=10, warmup_epochs=0)
    assert len(s2) == 30 and s2[0] == pytest.approx(1.0)
    # reference quirk (utils.py:525-538): warmup_steps without warmup_epochs shortens the table and trips its own assert
    with pytest.raises(AssertionError):
        utils.cosine_scheduler(1.0, 0.1, epochs=3, niter_per_ep=10, warmup_epochs=0, warmup_steps=5)


def test_moco_momentum_schedule():
    a = types.SimpleNamespace(epochs=10, moco_m=0.99)
    assert utils.adjust_moco_momentum(0, a) == pytest.approx(0.99)
    assert utils.adjust_moco_momentum(10, a) == pytest.approx(1.0)
    assert utils.adjust_moco_momentum(5, a) == pytest.approx(1 - 0.5 * (1 + math.cos(math.pi / 2)) * 0.01)


def test_meters():
    m = utils.SmoothedValue(window_size=3)
    for v in (1.0, 2.0, 3.0, 10.0):
        m.update(v)
    assert m.global_avg == pytest.approx(4.0) and m.median == 3.0 and m.max == 10.0 and m.value == 10.0
    ml = utils.MetricLogger()
    ml.update(a=1.0, b=torch.tensor(2.0), c=None)
    assert set(ml.meters) == {"a", "b"} and ml.b.global_avg == 2.0
    out = list(ml.log_every(range(3), 100, "hdr"))
    assert out == [0, 1, 2]


def test_scaler_state_has_scale():
    s = utils.NativeScalerWithGradNormCount()
    assert s.state_dict()["scale"] == 1.0          # engine logs it (E:157)


def test_engine_rejects_unbuilt_paths():
    from dig_b200 import ops
    from dig_b200.engine import train_one_epoch
    with pytest.raises(ops.DigError):          # only the patch4 models are built
        train_one_epoch(torch.nn.Linear(1, 1), None, None, [], None, None, "cpu", 0, None, normlize_target=False, patch_size=16,
                        args=types.SimpleNamespace(num_view=2))
    with pytest.raises(ops.DigError):          # the README runs two views
        train_one_epoch(torch.nn.Linear(1, 1), None, None, [], None, None, "cpu", 0, None, normlize_target=True, patch_size=4,
                        args=types.SimpleNamespace(num_view=1))


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception as e:  # report instead of letting the parent wait for its timeout
        q.put((rank, repr(e)))


def _worker_body(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dig_b200 import dist_layout
    from oracle import restatement as R
    torch.manual_seed(0)
    Q, C, T = 6, 16, 0.2
    q_all = torch.randn(world, 2 * Q, C)          # [q1 ; q2] per rank
    k_all = torch.nn.functional.normalize(torch.randn(world, 2 * Q, C), dim=-1)
    k1, k2 = dist_layout.gather_keys(k_all[rank].clone())     # [2, W*Q, C]: k1 of every rank, then k2 of every rank
    # rank-ordered concatenation, exactly torch.cat(all_gather(...)) of the reference
    ok = torch.equal(k1, k_all[:, :Q].reshape(world * Q, C)) and torch.equal(k2, k_all[:, Q:].reshape(world * Q, C))
    l1, a1, _ = R.contrastive_loss(q_all[rank, :Q], k2, T, rank)
    # single-process equivalent: all queries against all keys with global labels; per-rank loss is the mean over its own rows
    logits = torch.nn.functional.normalize(q_all[:, :Q].reshape(world * Q, C), dim=1) @ k2.t() / T
    labels = torch.arange(world * Q)
    ce = torch.nn.functional.cross_entropy(logits, labels, reduction="none") * (2 * T)
    ok = ok and abs(float(l1) - float(ce[rank * Q:(rank + 1) * Q].mean())) < 1e-5
    ok = ok and dist_layout.label_offset(Q, rank) == rank * Q
    stats = torch.tensor([1.0 + rank, 2.0])
    cnt = dist_layout.sync_batch_stats(stats, 10)
    ok = ok and cnt == 10.0 * world and float(stats[0]) == sum(1.0 + r for r in range(world)) and float(stats[1]) == 2.0 * world
    m = utils.SmoothedValue()
    m.update(float(rank + 1), n=1)
    m.synchronize_between_processes()
    ok = ok and m.global_avg == pytest.approx(sum(r + 1 for r in range(world)) / world)
    # DigDataParallel (dig_b200/parallel.py): construction broadcasts rank 0's parameters and buffers; the handle it leaves on the
    # wrapped module must not make the wrapper a child of its own child (train()/state_dict() would recurse)
    import dig_b200
    from dig_b200 import modeling  # noqa: F401
    from dig_b200.parallel import DigDataParallel
    torch.manual_seed(100 + rank)       # different initial weights per rank
    net = dig_b200.create_model("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128", pretrained=False, drop_path_rate=0.0,
                                drop_block_rate=None, mlp_dim=64, dim=32, T=0.2, num_windows=4, encoder_type="vit", queue_size=8,
                                patchnet_name="no_patchtrans")
    wrapped = DigDataParallel(net)
    wrapped.train()
    chk = torch.stack([p.detach().double().sum() for p in net.parameters()]).sum().reshape(1)
    both = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    ok = ok and all(torch.equal(b, both[0]) for b in both)
    ok = ok and net.__dict__.get("_dig_grad_sync") is wrapped and all(k.startswith("module.") for k in wrapped.state_dict())
    ok = ok and sum(1 for _ in wrapped.modules()) == 1 + sum(1 for _ in net.modules())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_layout_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_grad_segments_cover_the_flat_buffer_in_backward_order():
    """dig_b200.parallel.grad_segments: heads / 12 blocks / embed, contiguous, covering [0, total) (DigDataParallel all-reduces them)."""
    import dig_b200
    from dig_b200 import modeling  # noqa: F401
    from dig_b200.parallel import grad_segments
    m = dig_b200.create_model("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128", pretrained=False, drop_path_rate=0.0, drop_block_rate=None,
                              mlp_dim=256, dim=64, T=0.2, num_windows=4, encoder_type="vit", queue_size=8, patchnet_name="no_patchtrans")
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    named = dict(m.named_parameters())
    off, total = {}, 0
    for n in names:
        off[n] = total
        total += (named[n].numel() + 3) // 4 * 4
    seg = grad_segments(names, off, total)
    assert set(seg) == {"heads", "embed"} | {"block%d" % i for i in range(12)}
    spans = sorted(seg.values())
    assert spans[0][0] == 0 and spans[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert seg["embed"][0] == 0 and seg["heads"][1] == total
    assert seg["block3"] == (off["encoder.blocks.3.norm1.weight"], off["encoder.blocks.4.norm1.weight"])


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-batch", "2"], capture_output=True, text=True, timeout=600, cwd=root,
                         env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "crops/s" and line["higher_is_better"] is True and line["value"] > 0
    from oracle import ref_shims
    # the unmodified reference when /root/reference or the staged baseline/_ref exists, the oracle port otherwise
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_shims.reference_available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    assert line["e2e"] == {"value": line["value"], "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["metric"].startswith("pretrain text-crops/sec")


def test_bench_reference_arm_under_torchrun_rank0_prints_and_the_others_exit():
    """The driver launches the reference arm like the GPU arm (torchrun for N > 1): rank 0 alone runs it, and its one-rank process group
    must not go through the elastic agent's store (a tcp:// rendezvous hung there)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(29800 + os.getpid() % 100), os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0", "--cpu-batch", "2"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0


def test_peer_gradient_exchanges_cover_the_flat_buffer_once(monkeypatch):
    """PretrainStep._allreduce_segment / _peer_flush on the peer-memory path: the segments the backward finalises (heads, blocks 11..0,
    embed) are averaged in k overlapped exchanges (small co-resident blocks) plus one full-width exchange at the end, each a contiguous,
    16-byte aligned range, together covering [0, total) exactly once, with consecutive epochs on the gradient channel."""
    from dig_b200 import peer, pretrain_step

    calls = []
    monkeypatch.setattr(pretrain_step, "call", lambda name, *a: calls.append((name,) + a))

    class Comm:
        bases, world, rank = "bases", 8, 3

        def __init__(self):
            self.e = 0

        def next_epoch(self, ch):
            assert ch == peer.CH_GRADS
            self.e += 1
            return self.e

    for k, want_mid in ((3, 3), (2, 2), (0, 0)):
        calls.clear()
        seg, off = {}, 0
        for name, n in [("embed", 1000)] + [("block%d" % i, 400 + 8 * i) for i in range(12)] + [("heads", 2000)]:
            seg[name] = (off, off + n)
            off += n
        st = pretrain_step.PretrainStep.__new__(pretrain_step.PretrainStep)
        st.grad_seg, st._peer, st._grad_peer, st._peer_pending = seg, Comm(), "table", []
        st._peer_flush_keys = {"block%d" % (12 - (j * 12) // (k + 1)) for j in range(1, k + 1)}
        for key in ["heads"] + ["block%d" % i for i in reversed(range(12))] + ["embed"]:
            st._peer_pending.append(key)
            if key in st._peer_flush_keys:
                st._peer_flush(final=False)
        st._peer_flush(final=True)
        assert all(c[0] == "dig_peer_grad_allreduce" for c in calls)
        spans = sorted((c[7], c[7] + c[8]) for c in calls)
        assert spans[0][0] == 0 and spans[-1][1] == off and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(c[7] % 4 == 0 and c[8] % 4 == 0 for c in calls)
        assert [c[6] for c in calls] == list(range(1, len(calls) + 1))                 # epochs in issue order
        assert sum(1 for c in calls if c[10] == 1) == want_mid and calls[-1][10] == 0   # small blocks in the middle, full width at the end
        assert len(calls) == want_mid + 1


def test_device_prefetch_passes_batches_through_in_order():
    """engine._DevicePrefetch (CPU: no side stream): same items, same order, same length as the wrapped loader."""
    import torch
    from dig_b200.engine import _DevicePrefetch
    items = [((torch.full((2,), float(i)), torch.zeros(1), torch.ones(1)), "text%d" % i, i) for i in range(5)]
    feed = _DevicePrefetch(items, torch.device("cpu"))
    assert len(feed) == 5
    got = list(feed)
    assert [g[1] for g in got] == ["text%d" % i for i in range(5)]
    assert all(torch.equal(g[0][0], it[0][0]) for g, it in zip(got, items))
    assert list(_DevicePrefetch([], torch.device("cpu"))) == []
