"""Rank-layout helpers of the data-parallel step (host logic, device-agnostic so the gloo tests can drive them).

Reference semantics: `concat_all_gather` (modeling_pretrain_moco_mim_ori.py:580-591) concatenates the per-rank key
blocks in rank order, and `contrastive_loss` labels row i of rank r with `i + N*r` (M:453).  dig_b200 gathers
[k1 ; k2] of every rank with ONE all_gather_into_tensor and regroups it here.
"""
import torch
import torch.distributed as dist


def world_and_rank():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def gather_keys(kn, out_all=None, out2=None):
    """kn: [2Q, C] = normalised [k1 ; k2] of this rank.  Returns k_all2 [2, W*Q, C] (contiguous): k_all2[0] = k1 of every rank in rank
    order, k_all2[1] = k2 of every rank -- the row order `concat_all_gather` produces for each of the two calls (M:551-552)."""
    world, _ = world_and_rank()
    R, C = kn.shape
    Q = R // 2
    if world == 1:
        return kn.view(2, Q, C)
    if out_all is None:
        out_all = torch.empty(world * R, C, dtype=kn.dtype, device=kn.device)
    out_all = out_all.view(world * R, C)
    dist.all_gather_into_tensor(out_all, kn.contiguous())
    if out2 is None:
        out2 = torch.empty(2, world * Q, C, dtype=kn.dtype, device=kn.device)
    out2.view(2, world, Q, C).copy_(out_all.view(world, 2, Q, C).permute(1, 0, 2, 3))
    return out2


def label_offset(num_queries, rank):
    """First positive-key index of this rank's queries (M:453)."""
    return num_queries * rank


def sync_batch_stats(stats, rows):
    """SyncBatchNorm statistics exchange: all-reduce [sum | sumsq] and return the global row count."""
    world, _ = world_and_rank()
    if world > 1:
        dist.all_reduce(stats)
    return float(rows * world)
