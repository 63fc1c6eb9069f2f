"""Data-parallel wrapper for the dig_b200 pre-training model (optional replacement for torch's DistributedDataParallel, R:391).

`DistributedDataParallel` works with the model (gradients are ordinary leaf `.grad`s), but it copies every one of the 183 gradient
tensors into its buckets and back (2 x 183 small copy kernels per step) and starts its all-reduce only when the single autograd node has
returned -- ≈ 2 ms per step at bs=128.  `DigDataParallel` instead averages the model's flat gradient buffer in place, in 14 NCCL
all-reduces (heads, 12 blocks, patch embed) that `PretrainStep.backward` issues from its side stream as soon as a segment is final, so
they overlap the rest of the backward.  Same semantics as DDP for this model: parameters and buffers are broadcast from rank 0 once at
construction (afterwards every rank applies the same averaged gradients; BatchNorm buffers stay identical because SyncBatchNorm
statistics are global), gradients are averaged over the ranks.

    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    model = dig_b200.parallel.DigDataParallel(model)          # instead of DistributedDataParallel(model, device_ids=[gpu], ...)
"""
import torch
import torch.distributed as dist


class DigDataParallel(torch.nn.Module):
    def __init__(self, module, process_group=None):
        super().__init__()
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("DigDataParallel needs an initialised torch.distributed process group")
        self.module = module
        self.process_group = process_group
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)
        from . import ops
        ops.note_raw_parameter_write()      # written through .data: the bf16 operand shadows must be re-cast
        # PretrainStep.backward looks this handle up and averages its flat gradient buffer segment by segment.  It goes into the instance
        # __dict__ directly: a plain attribute assignment would register the wrapper as a child module of its own child (a cycle).
        module.__dict__["_dig_grad_sync"] = self

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def grad_segments(train_names, grad_off, grad_total):
    """[(name, start, end)] over the flat gradient buffer, in the order the backward finalises them: heads first, then encoder blocks from
    the last to the first, then the patch embedding.  Segments are contiguous because parameters are laid out in named_parameters order."""
    def group(n):
        if n.startswith("encoder.blocks."):
            return "block%d" % int(n.split(".")[2])
        return "embed" if n.startswith("encoder.") else "heads"
    bounds, cur = [], None
    for n in train_names:
        g = group(n)
        if g != cur:
            bounds.append([g, grad_off[n], None])
            if len(bounds) > 1:
                bounds[-2][2] = grad_off[n]
            cur = g
    bounds[-1][2] = grad_total
    names = [b[0] for b in bounds]
    if len(set(names)) != len(names):
        raise RuntimeError("trainable parameters are not grouped contiguously: %s" % names)
    return {b[0]: (b[1], b[2]) for b in bounds}
