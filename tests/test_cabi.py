"""C-ABI boundary (no GPU needed): the in-tree library builds, loads and exports every symbol include/dig_b200.h declares."""
import ctypes
import os

import pytest


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from dig_b200 import ops
    return ops.load()


def test_header_declares_entry_points():
    from dig_b200 import ops
    decls = ops.parse_header()
    for name in ("dig_gemm", "dig_attention_fwd", "dig_attention_bwd", "dig_layernorm_fwd", "dig_layernorm_bwd", "dig_mt_adamw",
                 "dig_mt_ema", "dig_infonce_rows", "dig_masked_mse", "dig_bn_apply", "dig_version", "dig_last_error"):
        assert name in decls
    assert len(decls) >= 30


def test_library_exports_every_declared_symbol(lib):
    from dig_b200 import ops
    for name in ops.parse_header():
        assert hasattr(lib, name), name


def test_library_links_no_driver(lib):
    """The .so must load on a box without libcuda (TMA encoder is resolved at run time)."""
    from dig_b200 import ops
    import subprocess
    out = subprocess.run(["ldd", ops.lib_path()], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out


def test_version_and_error_string(lib):
    assert lib.dig_version() == 3
    assert isinstance(lib.dig_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    """Bad descriptors are rejected before any CUDA call, with a message."""
    from dig_b200.ops import _Gemm
    assert lib.dig_gemm(None, None) == -1
    assert b"null descriptor" in lib.dig_last_error()
    g = _Gemm()
    g.M, g.N, g.K = 0, 128, 64
    assert lib.dig_gemm(ctypes.byref(g), None) == -1
    assert b"empty problem" in lib.dig_last_error()
    assert lib.dig_layernorm_fwd(None, None, None, None, None, None, 4, 100, 1e-6, 0, None) == -1
    assert lib.dig_infonce_rows(None, 0, 0, 0, 0.2, None, None) == -1


def test_cpu_tensors_are_refused(lib):
    import torch
    from dig_b200 import ops
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    with pytest.raises(ops.DigError):
        ops.gemm(a, a, torch.zeros(128, 128))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from dig_b200 import ops
    monkeypatch.setattr(ops, "_lib", None)
    monkeypatch.setattr(ops, "lib_path", lambda: str(tmp_path / "nope.so"))
    with pytest.raises(ops.DigError):
        ops.load()
