"""Host side of the NVLink peer-memory exchanges (dig_b200/csrc/peer.cu): one workspace per rank, IPC handles exchanged once through
torch.distributed, then the SyncBatchNorm statistics (R:390) and the MoCo key all-gather (M:580-591) of every step are carried by
dig_b200's own kernels -- no NCCL launch on the head chains.  torch.distributed remains the plumbing (rendezvous, the one-time handle
exchange) and still carries the gradient all-reduce.

Channels (one exchange sequence per stream, identical on every rank):
  0  online forward  BatchNorm statistics      (main stream)
  1  momentum forward BatchNorm statistics     (side stream)
  2  backward BatchNorm statistics             (main stream)
  3  key all-gather                            (side stream)
  4  gradient all-reduce                       (main stream, last kernel of the backward)
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import ops

CH_ONLINE, CH_MOMENTUM, CH_BACKWARD, CH_KEYS, CH_GRADS = 0, 1, 2, 3, 4
MAX_FLOATS = 8192
_MAX_RANKS = 8


class PeerComm:
    def __init__(self, device, key_table_bytes, group=None):
        lib = ops.load()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _MAX_RANKS:
            raise ops.DigError("peer-memory exchanges are built for one NVSwitch node (<= %d ranks)" % _MAX_RANKS)
        self.key_table_bytes = int(key_table_bytes)
        total, koff = ctypes.c_int64(), ctypes.c_int64()
        ops._check(lib.dig_peer_workspace_bytes(self.key_table_bytes, ctypes.byref(total), ctypes.byref(koff)), "dig_peer_workspace_bytes")
        self._own = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        ok, err = 1, ""
        try:
            ops._check(lib.dig_peer_alloc(total.value, ctypes.byref(self._own), handle), "dig_peer_alloc")
        except ops.DigError as e:
            ok, err = 0, str(e)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (ok, bytes(handle), torch.cuda.current_device(), os.uname().nodename), group=group)
        self._mapped = []
        bases = [0] * _MAX_RANKS
        if all(g[0] for g in gathered) and len({g[3] for g in gathered}) == 1:
            try:
                for r, g in enumerate(gathered):
                    if r == self.rank:
                        bases[r] = self._own.value
                    else:
                        p = ctypes.c_void_p()
                        ops._check(lib.dig_peer_open(g[1], ctypes.byref(p)), "dig_peer_open")
                        self._mapped.append(p)
                        bases[r] = p.value
            except ops.DigError as e:
                ok, err = 0, str(e)
        else:
            ok, err = 0, err or "a peer could not allocate its workspace, or the ranks span several hosts"
        flag = torch.tensor([ok], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)       # every rank maps every peer, or nobody uses the path
        self.ok = bool(int(flag.item()))
        self.error = err
        self.bases = (ctypes.c_int64 * _MAX_RANKS)(*bases)
        self.keys_offset = koff.value
        self.epoch = [0, 0, 0, 0, 0]
        self._group = group
        self._shared = []          # [(own pointer, [mapped peer pointers])] of alloc_shared
        if not self.ok:
            self.close()
        dist.barrier(group=group)

    def next_epoch(self, channel):
        self.epoch[channel] += 1
        return self.epoch[channel]

    def key_table_ptr(self, epoch):
        """Device address of this rank's key table for the exchange `epoch` ([2, W*Q, C] fp32, written by every rank's kernel)."""
        return self._own.value + self.keys_offset + (epoch & 1) * self.key_table_bytes

    def alloc_shared(self, nbytes):
        """COLLECTIVE: every rank allocates `nbytes` of device memory and maps every peer's allocation.  Returns (tensor-compatible own
        address, ctypes int64[8] table of all ranks' addresses) or None when a rank failed (then nobody uses it)."""
        lib = ops.load()
        own = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        ok = 1
        try:
            ops._check(lib.dig_peer_alloc(int(nbytes), ctypes.byref(own), handle), "dig_peer_alloc")
        except ops.DigError:
            ok = 0
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (ok, bytes(handle)), group=self._group)
        mapped, bases = [], [0] * _MAX_RANKS
        if all(g[0] for g in gathered):
            try:
                for r, g in enumerate(gathered):
                    if r == self.rank:
                        bases[r] = own.value
                    else:
                        p = ctypes.c_void_p()
                        ops._check(lib.dig_peer_open(g[1], ctypes.byref(p)), "dig_peer_open")
                        mapped.append(p)
                        bases[r] = p.value
            except ops.DigError:
                ok = 0
        else:
            ok = 0
        flag = torch.tensor([ok], device="cuda", dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self._group)
        if not int(flag.item()):
            for p in mapped:
                lib.dig_peer_close(p)
            if own:
                lib.dig_peer_free(own)
            return None
        self._shared.append((own, mapped))
        return own.value, (ctypes.c_int64 * _MAX_RANKS)(*bases)

    def check(self):
        e = ctypes.c_int32()
        ops._check(ops.load().dig_peer_error(self.bases, self.world, self.rank, ctypes.byref(e)), "dig_peer_error")
        if e.value:
            raise ops.DigError("a peer-memory exchange timed out: a rank never arrived")

    def close(self):
        lib = ops.load()
        for p in self._mapped:
            lib.dig_peer_close(p)
        self._mapped = []
        for own, mapped in getattr(self, "_shared", []):
            for p in mapped:
                lib.dig_peer_close(p)
            lib.dig_peer_free(own)
        self._shared = []
        if self._own:
            lib.dig_peer_free(self._own)
            self._own = ctypes.c_void_p()


class _DeviceArray:
    """Zero-copy torch view of raw device memory (torch.as_tensor understands __cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def shared_float_buffer(comm, n):
    """COLLECTIVE: a zero-initialised fp32 [n] tensor per rank in IPC-mapped memory + the table of every rank's address (or None)."""
    r = comm.alloc_shared(4 * int(n))
    if r is None:
        return None
    ptr, table = r
    t = torch.as_tensor(_DeviceArray(ptr, n), device=torch.device("cuda", torch.cuda.current_device()))
    return t, table


_comm = {}


def get(device, key_table_bytes):
    """The process-wide PeerComm (created collectively on first use; None when disabled with DIG_PEER=0, when the world is one rank, or
    when the workspaces could not be mapped -- the callers then fall back to torch.distributed collectives)."""
    if os.environ.get("DIG_PEER", "1") == "0" or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    if dist.get_backend() != "nccl":
        return None
    c = _comm.get("default")
    if c is None or (c.ok and c.key_table_bytes < key_table_bytes):
        keep = []
        if c is not None:
            torch.cuda.synchronize()
            dist.barrier()
            # a larger key table means a new workspace (collective); the shared gradient buffers handed out by alloc_shared stay mapped --
            # models hold tensors on them -- and move to the new communicator (flags and epochs restart from zero on every rank alike)
            keep, c._shared = c._shared, []
            c.close()
        c = PeerComm(device, key_table_bytes)
        c._shared.extend(keep)
        _comm["default"] = c
        if not c.ok and dist.get_rank() == 0:
            print("dig_b200: NVLink peer-memory exchanges unavailable (%s); using torch.distributed collectives" % c.error)
    return c if c.ok else None
