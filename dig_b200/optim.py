"""Fused multi-tensor AdamW for the pre-training step.

Same update rule and state layout as the reference's `custom_optim.AdamW`
(custom_optim/adamw.py:63-132, custom_optim/_functional.py:115-140: decoupled weight decay
`p *= 1 - lr*wd`, bias-corrected moments, eps added after `sqrt(v)/sqrt(bc2)`), same `param_groups`
protocol (`lr`, `weight_decay`, `betas`, `eps`, plus the engine's `lr_scale`), but ONE kernel launch
(dig_mt_adamw) over a device pointer table instead of ~8 ATen launches per parameter tensor.
Folded into the same launch: gradient unscale and clip_grad_norm_ (`grad_scale`, `max_norm`), the squared
gradient norm the engine logs (`sumsq_out`), and the refresh of the bf16 GEMM-operand shadows that
`PretrainStep` registered for the weights (`ops.register_shadow`).
"""
import torch

from . import ops
from .pretrain_step import MtTable

_HYPER_SLOTS = 4


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("amsgrad is not built (the reference runs with amsgrad=False)")
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))
        self._table = None
        self._sig = None
        self._hyper_sig = None
        self._pending = (1.0, None, None, None, None)   # (grad_scale, sumsq, max_norm, sumsq_out, guard) set by the loss scaler for the next step

    def set_grad_transform(self, grad_scale=1.0, sumsq=None, max_norm=None, sumsq_out=None, guard=None):
        """For the next step(): multiply gradients by `grad_scale`; clip them to `max_norm` (None = off) given `sumsq` (device fp32[1],
        squared norm of the unscaled gradients); accumulate the squared norm of the scaled gradients into `sumsq_out` (device fp32[1]);
        skip the whole update when the device scalar `guard` (the loss) is not finite."""
        self._pending = (float(grad_scale), sumsq, None if max_norm is None else float(max_norm), sumsq_out, guard)

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._sig = self._table = self._hyper_sig = None      # the moment tensors were replaced: rebuild the pointer table

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        self._sig = self._table = self._hyper_sig = None

    def _build(self, entries):
        dev = entries[0][0].device
        ps = [p.data for p, _ in entries]
        self._table = MtTable(dev, ps, [p.grad for p, _ in entries], [self.state[p]["exp_avg"] for p, _ in entries],
                              [self.state[p]["exp_avg_sq"] for p, _ in entries], [ops.shadow_of(p) for p, _ in entries])
        # per-tensor lr / weight decay staging: a small ring of pinned slots, each guarded by the event of the copy that last read it (the
        # engine lets the host run ahead of the GPU, so a single slot could be rewritten before its async H2D copy has executed)
        self._hyper_host = [torch.empty(2, len(entries), dtype=torch.float32).pin_memory() for _ in range(_HYPER_SLOTS)]
        self._hyper_event = [None] * _HYPER_SLOTS
        self._hyper_next = 0
        self._hyper_dev = torch.empty(2, len(entries), dtype=torch.float32, device=dev)
        self._hyper_sig = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        entries = [(p, g) for g in self.param_groups for p in g["params"] if p.grad is not None]
        if not entries:
            return loss
        if not entries[0][0].is_cuda:
            raise ops.DigError("FusedAdamW runs on CUDA tensors only (dig_b200 has no CPU path)")
        for p, _ in entries:
            st = self.state[p]
            if len(st) == 0:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        sig = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(),
                     0 if ops.shadow_of(p) is None else ops.shadow_of(p).data_ptr()) for p, _ in entries)
        if sig != self._sig:
            self._build(entries)
            self._sig = sig
        beta1, beta2 = entries[0][1]["betas"]
        eps = entries[0][1]["eps"]
        steps = set()
        lrs, wds = [], []
        for p, g in entries:
            if g["betas"] != (beta1, beta2) or g["eps"] != eps:
                raise ops.DigError("FusedAdamW needs the same betas/eps in every param group")
            lrs.append(g["lr"])
            wds.append(g["weight_decay"])
            st = self.state[p]
            st["step"] += 1
            steps.add(int(st["step"]))
        if len(steps) != 1:
            raise ops.DigError("FusedAdamW needs every parameter at the same step count")
        hyper = (tuple(lrs), tuple(wds))
        if hyper != self._hyper_sig:     # per-tensor lr / weight decay: one pinned staging write + one async H2D, only when they change
            i = self._hyper_next
            self._hyper_next = (i + 1) % _HYPER_SLOTS
            if self._hyper_event[i] is not None:
                self._hyper_event[i].synchronize()
            self._hyper_host[i].copy_(torch.tensor([lrs, wds], dtype=torch.float32))
            self._hyper_dev.copy_(self._hyper_host[i], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._hyper_event[i] = ev
            self._hyper_sig = hyper
        gs, sumsq, max_norm, sumsq_out, guard = self._pending
        if max_norm is not None and sumsq is None:
            raise ops.DigError("FusedAdamW: clipping needs the squared gradient norm (set_grad_transform(sumsq=...))")
        t = self._table
        lib = ops.load()
        rc = lib.dig_mt_adamw(t.ptrs[0].data_ptr(), t.ptrs[1].data_ptr(), t.ptrs[2].data_ptr(), t.ptrs[3].data_ptr(), t.ptrs[4].data_ptr(),
                              t.numel.data_ptr(), self._hyper_dev[0].data_ptr(), self._hyper_dev[1].data_ptr(),
                              t.blk_tensor.data_ptr(), t.blk_chunk.data_ptr(), t.num_blocks, beta1, beta2, eps, steps.pop(), gs,
                              None if sumsq is None else sumsq.data_ptr(), -1.0 if max_norm is None else max_norm,
                              None if sumsq_out is None else sumsq_out.data_ptr(), None if guard is None else guard.data_ptr(),
                              torch.cuda.current_stream().cuda_stream)
        ops.count_launch()
        if rc != 0:
            raise ops.DigError("dig_mt_adamw failed: %s" % lib.dig_last_error().decode())
        self._pending = (1.0, None, None, None, None)
        return loss


# the reference exposes the class as custom_optim.AdamW
AdamW = FusedAdamW
