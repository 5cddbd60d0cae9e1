"""CPU-only: the C-ABI library loads and exports every symbol include/mrfp_b200.h declares."""
import os
import re

from tests.common import ROOT


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mrfp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mrfp_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from mrfp_b200 import build, _lib
    build.build()
    lib = _lib.load()
    names = _declared_symbols()
    assert "mrfp_npplus_fwd_f32" in names and "mrfp_hrfp_fwd" in names
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SIGNATURES), (set(names) ^ set(_lib.SIGNATURES))


def test_strerror_and_version_without_gpu():
    from mrfp_b200 import _lib
    lib = _lib.load()
    assert lib.mrfp_version() >= 100
    assert lib.mrfp_strerror(0) == b"success"
    assert b"shape" in lib.mrfp_strerror(-2)
    assert lib.mrfp_npplus_ws_bytes(8, 256, 36864) >= 8 * 256 * 5 * 8
    assert lib.mrfp_npplus_ws_bytes(0, 1, 1) == 0


def test_host_wrappers_refuse_cpu_tensors_without_gpu():
    """No CPU fallback behind the Python host either: the autograd Functions raise before any launch."""
    import pytest
    import torch
    from mrfp_b200 import _lib
    from mrfp_b200.instnorm import instance_norm_relu
    from mrfp_b200.npplus import np_plus_with_draws, relu_with_plane_sums
    x = torch.randn(2, 8, 4, 4)
    with pytest.raises(_lib.MrfpError):
        instance_norm_relu(x)
    with pytest.raises(_lib.MrfpError):
        relu_with_plane_sums(x)
    with pytest.raises(_lib.MrfpError):
        np_plus_with_draws(x, torch.ones(2, 8), torch.zeros(2, 8))


def test_instnorm_argument_errors_without_gpu():
    """Argument validation of the InstanceNorm entry points happens before any CUDA call."""
    from mrfp_b200 import _lib
    lib = _lib.load()
    assert lib.mrfp_instnorm_fwd_f32(None, None, None, None, None, None, None, 1, 1, 1, 1e-5, 1, None) == -1
    assert lib.mrfp_instnorm_bwd_f32(None, None, None, None, None, None, None, None, None, 1, 1, 1, 1, None) == -1
    assert lib.mrfp_instnorm_fwd_f32(8, None, None, 8, 8, 8, None, 0, 1, 1, 1e-5, 1, None) == -2


def test_plan_fusion_bits_without_gpu():
    """mrfp_hrfp_plan_set_fusion is host-only state: bits 0-1 for a bf16 plan, always 0 for the other math modes."""
    import ctypes
    from mrfp_b200 import _lib
    lib = _lib.load()
    for mode, expect in ((2, (3, 1, 2, 0, 3)), (1, (0, 0, 0, 0, 0)), (0, (0, 0, 0, 0, 0))):
        h = ctypes.c_void_p()
        assert lib.mrfp_hrfp_plan_create(ctypes.byref(h), 2, 64, 12, 12, 48, 48, None, mode) == 0
        got = tuple(lib.mrfp_hrfp_plan_set_fusion(h, bits) for bits in (7, 1, 2, 0, 3))
        assert got == expect, (mode, got)
        lib.mrfp_hrfp_plan_destroy(h)
    assert lib.mrfp_hrfp_plan_set_fusion(None, 1) == -5
