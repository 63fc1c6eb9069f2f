"""ctypes binding of the dig_b200 C-ABI (include/dig_b200.h).

PyTorch is used only to own device memory and streams: every function here takes torch CUDA tensors,
passes their `data_ptr()` and sizes through the C-ABI and enqueues hand-written sm_100a kernels on the
current stream.  There is no fallback: a missing library or a failing call raises.
"""
import ctypes
import os

import torch

_LIB_NAME = "libdig_b200.so"
_lib = None


class DigError(RuntimeError):
    pass


def lib_path():
    """In-tree library; DIG_B200_LIB names an alternative build (A/B experiments, scripts/build_variant.py)."""
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), os.environ.get("DIG_B200_LIB", _LIB_NAME))


class _Gemm(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int64), ("N", ctypes.c_int64), ("K", ctypes.c_int64),
        ("A", ctypes.c_void_p), ("lda", ctypes.c_int64), ("a_mn_major", ctypes.c_int32),
        ("B", ctypes.c_void_p), ("ldb", ctypes.c_int64), ("b_mn_major", ctypes.c_int32),
        ("out", ctypes.c_void_p), ("ldo", ctypes.c_int64), ("out_fp32", ctypes.c_int32),
        ("bias", ctypes.c_void_p),
        ("residual", ctypes.c_void_p), ("ldr", ctypes.c_int64),
        ("res_row_mod", ctypes.c_int64),
        ("row_mask", ctypes.c_void_p), ("row_mask_value", ctypes.c_void_p),
        ("epilogue", ctypes.c_int32),
        ("aux", ctypes.c_void_p), ("ldaux", ctypes.c_int64),
        ("alpha", ctypes.c_float),
        ("split_k", ctypes.c_int32),
        ("colsum", ctypes.c_void_p),
        ("rowdot", ctypes.c_void_p), ("ldrowdot", ctypes.c_int64),
        ("aux_q8", ctypes.c_int32),
    ]


EPI_LINEAR, EPI_GELU, EPI_GELU_BWD, EPI_RELU_MASK, EPI_ROWDOT = 0, 1, 2, 3, 5


_CTYPE = {"double": ctypes.c_double, "int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "float": ctypes.c_float, "int": ctypes.c_int}


def header_path():
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "dig_b200.h")


def parse_header(path=None):
    """{function name: [ctypes argument types]} for every `int dig_*(...)` declared in include/dig_b200.h."""
    import re
    text = open(path or header_path()).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int|const char\*)\s+(dig_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        argtypes = []
        if args not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    argtypes.append(_CTYPE[a.replace("const ", "").split()[0]])
        decls[name] = (ret, argtypes)
    return decls


def load():
    """Load the in-tree shared library; raises DigError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.isfile(path):
        raise DigError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(dig_b200 has no CPU / eager fallback)" % path)
    lib = ctypes.CDLL(path)
    for name, (ret, argtypes) in parse_header().items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = ctypes.c_char_p if ret != "int" else ctypes.c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


# bf16 GEMM-operand shadows of fp32 master parameters, keyed by the parameter's data_ptr(): PretrainStep registers them, FusedAdamW looks
# them up so that its single launch also refreshes them (no separate fp32 -> bf16 cast pass per step).
import weakref

_shadows = weakref.WeakValueDictionary()      # entries vanish with the PretrainStep that owns the shadow tensors


def register_shadow(param, shadow):
    _shadows[param.data_ptr()] = shadow


def shadow_of(param):
    return _shadows.get(param.data_ptr())


_raw_writes = 0


def note_raw_parameter_write():
    """Call after writing parameters behind autograd's back (through `.data` or a raw pointer) WITHOUT refreshing their registered
    shadows: the next forward re-casts them.  (In-place torch ops on the parameters bump `Tensor._version`, which is watched too.)"""
    global _raw_writes
    _raw_writes += 1


def raw_parameter_writes():
    return _raw_writes


_launches = 0
_gemm_prof = None   # None, or a list of (event0, event1, flops) while profiling


def launch_count():
    """Number of dig_b200 kernels launched through this binding so far (one per C-ABI compute call)."""
    return _launches


def count_launch(n=1):
    global _launches
    _launches += n


def profile_gemm(enable):
    """Per-launch CUDA-event timing of dig_gemm.  profile_gemm(True) starts; profile_gemm(False) -> (flops, ms, launches,
    algorithmic HBM bytes: every operand and result of each launch counted once)."""
    global _gemm_prof
    if enable:
        _gemm_prof = []
        return None
    torch.cuda.synchronize()
    rec, _gemm_prof = _gemm_prof or [], None
    return (sum(r[2] for r in rec), sum(r[0].elapsed_time(r[1]) for r in rec), len(rec), sum(r[3] for r in rec))


def call(name, *args):
    """Generic C-ABI call: torch tensors are passed by data_ptr(), None as NULL; the current stream is appended."""
    lib = load()
    conv = [a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args]
    rc = getattr(lib, name)(*conv, torch.cuda.current_stream().cuda_stream)
    count_launch()
    if rc != 0:
        raise DigError("%s failed (%d): %s" % (name, rc, lib.dig_last_error().decode()))


def _check(rc, what):
    if rc != 0:
        raise DigError("%s failed (%d): %s" % (what, rc, load().dig_last_error().decode()))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _req(t, dtype, name):
    if not t.is_cuda:
        raise DigError("%s must be a CUDA tensor (dig_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise DigError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if t.stride(-1) != 1:
        raise DigError("%s must have a contiguous last dimension" % name)


def gemm(a, b, out, *, a_mn_major=False, b_mn_major=False, bias=None, residual=None, res_row_mod=0, row_mask=None,
         row_mask_value=None, epilogue=EPI_LINEAR, aux=None, alpha=1.0, split_k=1, colsum=None, rowdot=None):
    """out[M,N] = epilogue(alpha * A.B^T) on tcgen05 (see include/dig_b200.h).

    a: bf16 [M,K] (K-major) or [K,M] (a_mn_major); b: bf16 [N,K] or [K,N] (b_mn_major);
    out: bf16 or fp32 [M,N].  All 2-D, last dim contiguous.
    """
    _req(a, torch.bfloat16, "a"); _req(b, torch.bfloat16, "b")
    if a_mn_major:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn_major:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    if K != Kb or tuple(out.shape) != (M, N):
        raise DigError("gemm shape mismatch: a%s b%s out%s" % (tuple(a.shape), tuple(b.shape), tuple(out.shape)))
    g = _Gemm()
    g.M, g.N, g.K = M, N, K
    g.A, g.lda, g.a_mn_major = a.data_ptr(), a.stride(0), int(a_mn_major)
    g.B, g.ldb, g.b_mn_major = b.data_ptr(), b.stride(0), int(b_mn_major)
    g.out, g.ldo, g.out_fp32 = out.data_ptr(), out.stride(0), int(out.dtype == torch.float32)
    if out.dtype not in (torch.float32, torch.bfloat16):
        raise DigError("out must be fp32 or bf16")
    if bias is not None:
        _req(bias, torch.float32, "bias")
        g.bias = bias.data_ptr()
    if residual is not None:
        _req(residual, torch.float32, "residual")
        g.residual, g.ldr = residual.data_ptr(), residual.stride(0)
    g.res_row_mod = res_row_mod
    if row_mask is not None:
        _req(row_mask, torch.uint8, "row_mask"); _req(row_mask_value, torch.float32, "row_mask_value")
        g.row_mask, g.row_mask_value = row_mask.data_ptr(), row_mask_value.data_ptr()
    g.epilogue = epilogue
    if aux is not None:
        if aux.dtype == torch.uint8:        # 8-bit GELU pre-activation codes (dig_gemm_t.aux_q8)
            if epilogue not in (EPI_GELU, EPI_GELU_BWD):
                raise DigError("a uint8 aux tensor (8-bit pre-activation codes) is for the GELU / GELU' epilogues only")
            _req(aux, torch.uint8, "aux")
            g.aux_q8 = 1
        else:
            _req(aux, torch.bfloat16, "aux")
        g.aux, g.ldaux = aux.data_ptr(), aux.stride(0)
    g.alpha = alpha
    g.split_k = split_k
    if colsum is not None:
        _req(colsum, torch.float32, "colsum")
        g.colsum = colsum.data_ptr()
    if rowdot is not None:
        _req(rowdot, torch.float32, "rowdot")
        if rowdot.dim() != 2 or rowdot.shape[0] != M or rowdot.shape[1] * 64 < N:
            raise DigError("rowdot must be fp32 [M, >= N/64], got %s" % (tuple(rowdot.shape),))
        g.rowdot, g.ldrowdot = rowdot.data_ptr(), rowdot.stride(0)
    if _gemm_prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _check(load().dig_gemm(ctypes.byref(g), _stream()), "dig_gemm")
        e1.record()
        osz = out.element_size()
        nbytes = 2.0 * (M * K + N * K) + M * N * osz * (split_k if split_k > 1 else 1)
        if residual is not None:
            nbytes += 4.0 * M * N
        if aux is not None:
            nbytes += float(aux.element_size()) * M * N
        _gemm_prof.append((e0, e1, 2.0 * M * N * K, nbytes))
    else:
        _check(load().dig_gemm(ctypes.byref(g), _stream()), "dig_gemm")
    count_launch()
    return out


_attn_prof = None   # None, or {"fwd": [(e0, e1, flops, exps)], "bwd": [...]} while profiling


def profile_attention(enable):
    """Per-launch CUDA-event timing of the fused attention kernels.  profile_attention(False) -> {"fwd"/"bwd": (flops, ms, launches,
    exp2 count)} -- algorithmic FLOPs 4*256*256*64 (forward) / 10*256*256*64 (backward) per (sequence, head), no recompute counted."""
    global _attn_prof
    if enable:
        _attn_prof = {"fwd": [], "bwd": []}
        return None
    torch.cuda.synchronize()
    rec, _attn_prof = _attn_prof or {"fwd": [], "bwd": []}, None
    return {k: (sum(r[2] for r in v), sum(r[0].elapsed_time(r[1]) for r in v), len(v), sum(r[3] for r in v)) for k, v in rec.items()}


def _attn_timed(kind, items, fn):
    if _attn_prof is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    per = 256 * 256 * 64
    _attn_prof[kind].append((e0, e1, float(items) * per * (4 if kind == "fwd" else 10), float(items) * 256 * 256 * (1 if kind == "fwd" else 1)))
    return r


def attention_fwd(qkv, out, lse, heads, scale, p_in_smem=False):
    """Fused softmax(q k^T * scale) v over 256-token sequences (F:97-118). qkv bf16 [S*256, 3*heads*64]."""
    _req(qkv, torch.bfloat16, "qkv"); _req(out, torch.bfloat16, "out")
    d = heads * 64
    rows = qkv.shape[0]
    if qkv.shape[1] != 3 * d or rows % 256 or tuple(out.shape) != (rows, d) or not qkv.is_contiguous() or not out.is_contiguous():
        raise DigError("attention_fwd: bad shapes qkv%s out%s heads=%d" % (tuple(qkv.shape), tuple(out.shape), heads))
    if lse is not None:
        _req(lse, torch.float32, "lse")
        if lse.numel() != rows // 256 * heads * 256:
            raise DigError("attention_fwd: lse must hold [S, heads, 256]")
    _attn_timed("fwd", rows // 256 * heads, lambda: _check(load().dig_attention_fwd(
        _ptr(qkv), _ptr(out), _ptr(lse), rows // 256, heads, scale, int(p_in_smem), _stream()), "dig_attention_fwd"))
    count_launch()
    return out


def attention_bwd(qkv, out, dout, lse, dqkv, heads, scale):
    """Gradient of attention_fwd w.r.t. qkv."""
    for t, n in ((qkv, "qkv"), (out, "out"), (dout, "dout"), (dqkv, "dqkv")):
        _req(t, torch.bfloat16, n)
        if not t.is_contiguous():
            raise DigError("attention_bwd: %s must be contiguous" % n)
    _req(lse, torch.float32, "lse")
    d = heads * 64
    rows = qkv.shape[0]
    if qkv.shape != dqkv.shape or out.shape != dout.shape or qkv.shape[1] != 3 * d or tuple(out.shape) != (rows, d) or rows % 256:
        raise DigError("attention_bwd: bad shapes")
    _check(load().dig_attention_bwd(_ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv), rows // 256, heads, scale, _stream()),
           "dig_attention_bwd")
    count_launch()
    return dqkv


def attention_bwd_d(qkv, dout, lse, dsum, dqkv, heads, scale):
    """Gradient of attention_fwd w.r.t. qkv with D = rowsum(dout o out) per (token, head) precomputed (fp32 [S*256, heads]; `gemm(...,
    epilogue=EPI_ROWDOT, aux=out, rowdot=dsum)` on the output-projection dgrad emits it): persistent kernel, `out` is not read."""
    for t, n in ((qkv, "qkv"), (dout, "dout"), (dqkv, "dqkv")):
        _req(t, torch.bfloat16, n)
        if not t.is_contiguous():
            raise DigError("attention_bwd_d: %s must be contiguous" % n)
    _req(lse, torch.float32, "lse"); _req(dsum, torch.float32, "dsum")
    d = heads * 64
    rows = qkv.shape[0]
    if qkv.shape != dqkv.shape or qkv.shape[1] != 3 * d or tuple(dout.shape) != (rows, d) or rows % 256 or \
            tuple(dsum.shape) != (rows, heads) or not dsum.is_contiguous():
        raise DigError("attention_bwd_d: bad shapes")
    _attn_timed("bwd", rows // 256 * heads, lambda: _check(load().dig_attention_bwd_d(
        _ptr(qkv), _ptr(dout), _ptr(lse), _ptr(dsum), _ptr(dqkv), rows // 256, heads, scale, _stream()), "dig_attention_bwd_d"))
    count_launch()
    return dqkv
