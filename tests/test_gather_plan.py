"""Host logic of the operand-fused convolutions (csrc/conv_gather.cu): the shared-memory plan of a launch.  No device call."""
import ctypes

import numpy as np

from mrfp_b200 import _lib

SMEM_LIMIT = 232448          # 227 KB opt-in maximum per CTA on sm_100


def _plan(mode, cout, cin, dil, lo_h=None, lo_w=None, h=0, w=0):
    lib = _lib.load()
    fn = lib.mrfp_debug_gather_plan
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p] * 2 + [ctypes.c_int] * 2 + [ctypes.POINTER(ctypes.c_int)]
    out = (ctypes.c_int * 6)()
    a = None if lo_h is None else np.ascontiguousarray(lo_h, dtype=np.int32)
    b = None if lo_w is None else np.ascontiguousarray(lo_w, dtype=np.int32)
    rc = fn(mode, cout, cin, dil, None if a is None else a.ctypes.data, None if b is None else b.ctypes.data, h, w, out)
    assert rc == 0
    return dict(mt=out[0], nA=out[1], nB=out[2], smem=out[3], e_bytes=out[4], boxw=out[5])


def _lo(src, dst):
    """First replica of each source index under ATen's nearest rule (size= form), [src + 1]."""
    scale = np.float32(src) / np.float32(dst)
    idx = np.minimum(np.floor(np.arange(dst, dtype=np.float32) * scale).astype(np.int64), src - 1)
    return np.searchsorted(idx, np.arange(src + 1)).astype(np.int32)


# (cin, cout, dil) of the reference chain (deepv3.py:221-237)
LAYERS = [(64, 64, 1), (64, 64, 1), (64, 128, 2), (128, 256, 2), (256, 128, 1), (128, 64, 1), (64, 64, 2), (64, 64, 2)]


def test_forward_plans_fit_for_every_stage():
    for cin, cout, dil in LAYERS[1:]:
        p = _plan(0, cout, cin, dil)
        assert p["nA"] >= 2 and p["nB"] >= 4 and p["smem"] <= SMEM_LIMIT, (cin, cout, dil, p)
        assert p["mt"] == (1 if cout == 256 else 2) and p["boxw"] == 8 * p["mt"] + 2 * dil


def test_backward_plans_of_the_non_replicating_stages_fit():
    # dgrad of stage k: cout_k channels in, cin_k channels out
    for cin, cout, dil in LAYERS[4:]:
        p = _plan(1, cin, cout, dil)
        assert p["nA"] == 2 and p["nB"] >= 3 and p["smem"] <= SMEM_LIMIT, (cin, cout, dil, p)


def test_replica_plans_fit_only_where_the_staging_has_room():
    # 768^2 crop: conv resolutions 192 -> 231 -> 277 -> 332 -> 384
    res = [192, 231, 277, 332, 384]
    for k in range(4):
        cin, cout, dil = LAYERS[k]
        lo = _lo(res[k], res[k + 1])
        assert int((lo[1:] - lo[:-1]).max()) == 2
        p = _plan(2, cin, cout, dil, lo, lo, res[k], res[k])
        if k < 2:        # 64 -> 64, dilation 1: two halo stages + side stage + extras stage fit on the 16 x 16 tile
            assert p["nA"] == 2 and p["mt"] == 2 and p["smem"] <= SMEM_LIMIT, (k, p)
            # the extras stage holds every additional replica of the worst tile: at most (1.21^2 - 1) of the box and a margin
            npix = p["boxw"] * (16 + 2 * dil)
            assert 0 < p["e_bytes"] <= 128 * npix
        else:            # dilation 2 and 128 / 256 channels: refused, the chain keeps the separate apply pass there
            assert p["nA"] == 0, (k, p)


def test_replica_plan_needs_the_host_tables():
    assert _plan(2, 64, 64, 1)["nA"] == 0
