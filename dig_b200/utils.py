"""Host-side helpers the pre-training engine calls every step (reference utils/utils.py, tag U):
smoothed meters (U:30-282), the loss scaler protocol with gradient-norm (U:477-519), cosine schedules
(U:522-543).  Device work (gradient norm, unscale, clipping) runs in the multi-tensor kernels.
"""
import datetime
import math
import time
from collections import defaultdict, deque

import numpy as np
import torch
import torch.distributed as dist

from . import ops


def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process():
    return get_rank() == 0


class SmoothedValue(object):
    """Windowed median / average plus global average of one scalar series (U:30-93)."""

    def __init__(self, window_size=20, fmt=None):
        self.deque = deque(maxlen=window_size)
        self.total = 0.0
        self.count = 0
        self.fmt = fmt or "{avg:.4f} ({global_avg:.4f})"

    def update(self, value, n=1):
        self.deque.append(value)
        self.count += n
        self.total += value * n

    def synchronize_between_processes(self):
        if not is_dist_avail_and_initialized():
            return
        dev = "cuda" if torch.cuda.is_available() and dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([self.count, self.total], dtype=torch.float64, device=dev)
        dist.barrier()
        dist.all_reduce(t)
        t = t.tolist()
        self.count, self.total = int(t[0]), t[1]

    @property
    def median(self):
        return float(np.median(np.asarray(self.deque, dtype=np.float64)))

    @property
    def avg(self):
        return float(np.mean(np.asarray(self.deque, dtype=np.float64)))

    @property
    def global_avg(self):
        return self.total / self.count

    @property
    def max(self):
        return max(self.deque)

    @property
    def value(self):
        return self.deque[-1]

    def __str__(self):
        return self.fmt.format(median=self.median, avg=self.avg, global_avg=self.global_avg, max=self.max, value=self.value)


class MetricLogger(object):
    def __init__(self, delimiter="\t"):
        self.meters = defaultdict(SmoothedValue)
        self.delimiter = delimiter

    def update(self, **kwargs):
        for k, v in kwargs.items():
            if v is None:
                continue
            if isinstance(v, torch.Tensor):
                v = v.item()
            assert isinstance(v, (float, int))
            self.meters[k].update(v)

    def __getattr__(self, attr):
        if attr in self.meters:
            return self.meters[attr]
        if attr in self.__dict__:
            return self.__dict__[attr]
        raise AttributeError("'{}' object has no attribute '{}'".format(type(self).__name__, attr))

    def __str__(self):
        return self.delimiter.join("{}: {}".format(n, str(m)) for n, m in self.meters.items())

    def synchronize_between_processes(self):
        for meter in self.meters.values():
            meter.synchronize_between_processes()

    def add_meter(self, name, meter):
        self.meters[name] = meter

    def log_every(self, iterable, print_freq, header=None):
        header = header or ""
        start = end = time.time()
        iter_time, data_time = SmoothedValue(fmt="{avg:.4f}"), SmoothedValue(fmt="{avg:.4f}")
        n = len(iterable)
        width = str(len(str(n)))
        for i, obj in enumerate(iterable):
            data_time.update(time.time() - end)
            yield obj
            iter_time.update(time.time() - end)
            if i % print_freq == 0 or i == n - 1:
                eta = str(datetime.timedelta(seconds=int(iter_time.global_avg * (n - i))))
                parts = [header, ("[{0:" + width + "d}/{1}]").format(i, n), "eta: " + eta, str(self), "time: " + str(iter_time),
                         "data: " + str(data_time)]
                if torch.cuda.is_available():
                    parts.append("max mem: {:.0f}".format(torch.cuda.max_memory_allocated() / (1024.0 * 1024.0)))
                print(self.delimiter.join(parts))
            end = time.time()
        total = time.time() - start
        print("{} Total time: {} ({:.4f} s / it)".format(header, str(datetime.timedelta(seconds=int(total))), total / max(n, 1)))


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0, warmup_steps=-1):
    """U:522-538 (warmup_steps is only honoured when warmup_epochs > 0, as in the reference)."""
    warmup_iters = warmup_steps if warmup_steps > 0 else warmup_epochs * niter_per_ep
    print("Set warmup steps = %d" % warmup_iters)
    warm = np.linspace(start_warmup_value, base_value, warmup_iters) if warmup_epochs > 0 else np.array([])
    iters = np.arange(epochs * niter_per_ep - warmup_iters)
    sched = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * iters / len(iters)))
    sched = np.concatenate((warm, sched))
    assert len(sched) == epochs * niter_per_ep
    return sched


def adjust_moco_momentum(epoch, args):
    """U:540-543"""
    return 1.0 - 0.5 * (1.0 + math.cos(math.pi * epoch / args.epochs)) * (1.0 - args.moco_m)


class GradNorm:
    """Global L2 gradient norm (U:507-519) in one multi-tensor launch."""

    def __init__(self):
        self._sig, self._table, self._out = None, None, None

    def sumsq(self, parameters):
        from .pretrain_step import MtTable
        grads = [p.grad for p in parameters if p.grad is not None]
        if not grads:
            return None
        sig = tuple(g.data_ptr() for g in grads)
        if sig != self._sig:
            self._table = MtTable(grads[0].device, grads)
            self._out = torch.zeros(1, dtype=torch.float32, device=grads[0].device)
            self._sig = sig
        self._out.zero_()
        t = self._table
        lib = ops.load()
        rc = lib.dig_mt_sumsq(t.ptrs[0].data_ptr(), t.numel.data_ptr(), t.blk_tensor.data_ptr(), t.blk_chunk.data_ptr(), t.num_blocks,
                              self._out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ops.count_launch()
        if rc != 0:
            raise ops.DigError("dig_mt_sumsq failed: %s" % lib.dig_last_error().decode())
        return self._out


_grad_norm = GradNorm()


def get_grad_norm_(parameters, norm_type=2.0):
    if float(norm_type) != 2.0:
        raise NotImplementedError("only the L2 norm is built (the engine never asks for another)")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    s = _grad_norm.sumsq(list(parameters))
    return torch.tensor(0.0) if s is None else s.sqrt().reshape(())


class NativeScalerWithGradNormCount:
    """Same call protocol as U:477-504.  The B200 path computes in bf16 (fp32 range), so the loss scale is a constant 1.0:
    `state_dict()["scale"]` exists because the engine logs it (E:157)."""
    state_dict_key = "amp_scaler"

    def __init__(self):
        self._scale = 1.0
        self._gn = GradNorm()
        self._sumsq_out = None

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True):
        loss.backward(create_graph=create_graph)
        if not update_grad:
            return None
        assert parameters is not None
        params = [p for p in parameters]
        fused = hasattr(optimizer, "set_grad_transform")
        if clip_grad is not None:
            # U:487-490: `clip_grad is not None` clips -- also to 0.0, which zeroes every gradient (the engine's default max_norm=0)
            sumsq = self._gn.sumsq(params)
            norm = torch.tensor(0.0) if sumsq is None else sumsq.sqrt().reshape(())
            if fused:
                optimizer.set_grad_transform(1.0, sumsq, clip_grad, None, self._guard(loss))      # clipping folded into the AdamW launch
            else:
                torch.nn.utils.clip_grad_norm_(params, clip_grad)
            optimizer.step()
        elif fused and any(p.grad is not None for p in params):
            # U:492-493 get_grad_norm_: the AdamW launch sums the squared gradients it reads anyway (no second pass over them)
            if self._sumsq_out is None or self._sumsq_out.device != loss.device:
                self._sumsq_out = torch.zeros(1, dtype=torch.float32, device=loss.device)
            self._sumsq_out.zero_()
            optimizer.set_grad_transform(1.0, None, None, self._sumsq_out, self._guard(loss))
            optimizer.step()
            norm = self._sumsq_out.sqrt().reshape(())
        else:
            sumsq = self._gn.sumsq(params)
            norm = torch.tensor(0.0) if sumsq is None else sumsq.sqrt().reshape(())
            optimizer.step()
        return norm

    @staticmethod
    def _guard(loss):
        g = loss.detach()
        return g.reshape(1) if (g.is_cuda and g.dtype == torch.float32 and g.numel() == 1) else None

    def state_dict(self):
        return {"scale": self._scale}

    def load_state_dict(self, state_dict):
        self._scale = float(state_dict.get("scale", 1.0))
