"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`).

SURVEY.md section 4(iv): W ranks x B crops over NCCL/NVLink must equal ONE rank x (W*B) crops with the batch concatenated in rank order
-- SyncBatchNorm statistics are global, the keys of all ranks are the negatives, labels are offset by Q*rank (M:453), gradients are
averaged -- so the rank-averaged loss and the averaged gradients of the 2-rank step are compared with a single-process step at 2B, and
the NVLink peer-memory exchanges (csrc/peer.cu) with the torch.distributed collectives they replace (DIG_PEER=0).
"""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
MODEL = "pretrain_simmim_moco_ori_vit_tiny_patch4_32x128"
KW = dict(pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2, num_windows=4, encoder_type="vit",
          queue_size=65536, patchnet_name="no_patchtrans")
B = 16


def _model():
    import dig_b200
    from dig_b200 import modeling  # noqa: F401
    torch.manual_seed(0)
    return dig_b200.create_model(MODEL, **KW).train()


def _step(net, model, img, aug, mk):
    from dig_b200.engine import masked_pixel_mse
    out = net(img, aug, mk, 0.99, True)
    lpix = masked_pixel_mse(out["vis_out"][0], img, mk[:, 0])
    loss = out["contra_loss"] * 0.1 + lpix
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    stats = {k: v.detach().clone() for k, v in model.state_dict().items() if "running_" in k}
    return float(loss), float(out["contra_loss"]), float(lpix), grads, stats


def _worker(rank, world, port, q):
    try:
        import datetime
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        from oracle import restatement as R
        img, aug, mask = R.synthetic_batch(world * B, seed=3)
        mk = mask.clone()
        mk[:, 1, :] = False
        ref = None
        if rank == 0:
            # the single-process step over the concatenated batch, BEFORE the process group exists (with a group of two ranks the model
            # would -- correctly -- wait for the other rank's keys)
            import __graft_entry__ as ge
            ge.build()
            single = _model().to(dev)
            ref = _step(single, single, img.to(dev), aug.to(dev), mk.to(dev))
            del single
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=180))
        dist.barrier()
        from dig_b200.parallel import DigDataParallel
        sl = slice(rank * B, (rank + 1) * B)
        res = {}
        for tag, env in (("peer", "1"), ("nccl", "0")):
            os.environ["DIG_PEER"] = env
            model = _model().to(dev)
            net = DigDataParallel(torch.nn.SyncBatchNorm.convert_sync_batchnorm(model))
            res[tag] = _step(net, model, img[sl].to(dev), aug[sl].to(dev), mk[sl].to(dev))
            res[tag + "_peer_used"] = model._step._peer is not None
        # rank-averaged losses
        for tag in ("peer", "nccl"):
            t = torch.tensor(res[tag][:3], device=dev, dtype=torch.float64)
            dist.all_reduce(t)
            res[tag + "_avg"] = (t / world).tolist()
        out = None
        if rank == 0:
            def rel(a, b):
                return float((a - b).norm() / (b.norm() + 1e-20))
            out = {"peer_used": res["peer_peer_used"], "nccl_peer_used": res["nccl_peer_used"], "loss_single": ref[:3],
                   "loss_peer": res["peer_avg"], "loss_nccl": res["nccl_avg"],
                   "grad_rel_peer_vs_single": {n: rel(res["peer"][3][n], ref[3][n]) for n in ref[3]},
                   "grad_rel_peer_vs_nccl": {n: rel(res["peer"][3][n], res["nccl"][3][n]) for n in ref[3]},
                   "bn_absdiff_peer_vs_single": {k: float((res["peer"][4][k] - ref[4][k]).abs().max()) for k in ref[4]}}
        # gradients are identical on every rank after the averaging
        chk = torch.stack([g.double().sum() for g in res["peer"][3].values()]).sum().reshape(1)
        both = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        same = all(torch.equal(b, both[0]) for b in both)
        q.put((rank, same, out))
        dist.destroy_process_group()
    except Exception as e:      # report instead of letting the parent wait for its timeout
        import traceback
        q.put((rank, False, "%r\n%s" % (e, traceback.format_exc())))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_equal_one_rank_with_the_concatenated_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=400) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    assert all(r[1] for r in res), res
    out = res[0][2]
    print(out["loss_single"], out["loss_peer"], out["loss_nccl"])
    assert out["peer_used"] and not out["nccl_peer_used"]
    for got in (out["loss_peer"], out["loss_nccl"]):
        for a, b in zip(got, out["loss_single"]):
            assert a == pytest.approx(b, rel=1e-3)
    worst = max(out["grad_rel_peer_vs_single"].values())
    worst_pn = max(out["grad_rel_peer_vs_nccl"].values())
    print("worst grad rel-L2 vs single-process 2B step: %.3e; peer vs nccl: %.3e" % (worst, worst_pn))
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        import json
        json.dump(out, open(os.path.join(d, "multirank_parity.json"), "w"), indent=1)
    # run-to-run noise of the atomically accumulated statistics is ~1e-2 on head tensors at small batch
    assert worst < 8e-2 and worst_pn < 8e-2
    # BatchNorm running statistics (global over the ranks == over the concatenated batch); absolute, because several running means are
    # exact zeros up to rounding (the predictor's first Linear sees a zero-mean input: the projector ends in BatchNorm(affine=False))
    assert max(out["bn_absdiff_peer_vs_single"].values()) < 2e-3
