"""Operand-fused bf16 chain (csrc/conv_gather.cu) vs the same chain with the element-wise passes as separate kernels.

Both compute the same arithmetic on the same bf16-rounded operands — A_k = ReLU(BN(nearest(Y_{k-1}))) (forward, fusion bit 0)
and dY_k = BN-backward apply of (dA_{k+1}, Y_k) (backward of the non-replicating stages, bit 1) are rounded to bf16 once
in either path, into shared memory in one and into HBM in the other — so the paths may differ only by the fp32
accumulation order of the convolution (tap-major vs chunk-major), i.e. by isolated bf16 rounding flips that the later
stages carry along.  The unfused path is the one pinned
kernel by kernel in tests/test_hrfp_stage_gpu.py and tests/test_conv_tc_gpu.py; the fused path (both bits) is the default
of every other chain test (oracle, reference fixtures, full size)."""
import math

import numpy as np
import pytest
import torch

from oracle import mrfp_oracle as O
from tests.common import make_hrfp_params, make_feat
from tests.test_hrfp_gpu import _modules

pytestmark = pytest.mark.gpu


def _chain(xp_np, ws, gs, h, w, fuse, g1, g2):
    from mrfp_b200.hrfp import hrfp_chain, MATH_BF16
    convs, bns = _modules(ws, gs, "cuda")
    xp = torch.from_numpy(xp_np).cuda().requires_grad_(True)
    out, dec = hrfp_chain(xp, convs, bns, h, w, math_mode=MATH_BF16, fuse=fuse)
    torch.autograd.backward([out, dec], [g1, g2])
    rm = torch.stack([b.running_mean[:64] for b in bns])
    rv = torch.stack([b.running_var[:64] for b in bns])
    return out.detach(), dec.detach(), xp.grad.detach(), rm, rv


def _l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


# (n, h, w): odd sizes exercise ragged tile edges (16 x 16 and 16 x 8 tiles), 40 x 56 a non-square image, 200 x 184
# several tiles per image row and column, 16 x 20 an image smaller than one halo tile
GEOMS = [(2, 48, 48), (3, 40, 56), (2, 60, 44), (2, 96, 80), (2, 200, 184), (2, 16, 20)]


@pytest.mark.parametrize("fuse", [1, 2, 3])
@pytest.mark.parametrize("geom", GEOMS)
def test_fused_equals_unfused(geom, fuse):
    n, h, w = geom
    xh, xw = math.ceil(h / 4), math.ceil(w / 4)
    ws, gs = make_hrfp_params(11 + h)
    xp = make_feat(12 + w, (n, 64, xh, xw))
    gen = torch.Generator(device="cuda").manual_seed(5)
    g1 = torch.randn((n, 64, xh, xw), device="cuda", generator=gen)
    g2 = torch.randn((n, 256, h // 2, w // 2), device="cuda", generator=gen)
    ref = _chain(xp, ws, gs, h, w, 0, g1, g2)
    got = _chain(xp, ws, gs, h, w, fuse, g1, g2)
    names = ("OCout", "OCout_dec", "g_xp", "running_mean", "running_var")
    # forward: accumulation-order noise only (bf16 storage: a flipped rounding is 4e-3 of one element)
    tol = dict(OCout=8e-3, OCout_dec=8e-3, g_xp=6e-2, running_mean=1e-3, running_var=1e-3)
    for name, a, b in zip(names, got, ref):
        assert torch.isfinite(a).all(), name
        assert _l2(a, b) <= tol[name], (name, fuse, geom, _l2(a, b))


def test_fusion_bits_are_reported_and_ignored_outside_bf16():
    from mrfp_b200.hrfp import HrfpPlan, MATH_BF16, MATH_TF32
    assert HrfpPlan(2, 64, 12, 12, 48, 48, "cuda", MATH_BF16).fuse == 3
    assert HrfpPlan(2, 64, 12, 12, 48, 48, "cuda", MATH_BF16, fuse=1).fuse == 1
    assert HrfpPlan(2, 64, 12, 12, 48, 48, "cuda", MATH_BF16, fuse=0).fuse == 0
    assert HrfpPlan(2, 64, 12, 12, 48, 48, "cuda", MATH_TF32, fuse=3).fuse == 0


@pytest.mark.parametrize("fuse", [0, 3])
def test_vs_oracle_small(fuse):
    """Both paths against the numpy oracle (fp64 arithmetic) at the bf16 mode's stated tolerance."""
    n, h, w = 2, 48, 48
    ws, gs = make_hrfp_params(4)
    xp = make_feat(5, (n, 64, 12, 12))
    ws64 = [t.astype(np.float64) for t in ws]; gs64 = [t.astype(np.float64) for t in gs]
    ro, rd, saved = O.hrfp_forward(xp.astype(np.float64), ws64, gs64, h, w)
    g1 = np.random.default_rng(6).standard_normal(ro.shape).astype(np.float32)
    g2 = np.random.default_rng(7).standard_normal(rd.shape).astype(np.float32)
    rg = O.hrfp_backward(g1.astype(np.float64), g2.astype(np.float64), ws64, gs64, saved)
    out, dec, gx, _, _ = _chain(xp, ws, gs, h, w, fuse, torch.from_numpy(g1).cuda(), torch.from_numpy(g2).cuda())
    assert _l2(out.cpu(), torch.from_numpy(ro)) <= 3e-2
    assert _l2(dec.cpu(), torch.from_numpy(rd)) <= 3e-2
    assert _l2(gx.cpu(), torch.from_numpy(rg)) <= 3e-1


def test_fused_equals_unfused_on_a_sweep_of_geometries():
    """A deterministic sweep of odd image sizes (ragged tiles on every side, single-tile images, tall and wide images,
    batch 1): the operand-fused chain against the separate passes."""
    rng = np.random.default_rng(2024)
    geoms = [(int(rng.integers(1, 4)), int(rng.integers(16, 150)), int(rng.integers(16, 150))) for _ in range(10)]
    geoms += [(1, 17, 131), (2, 129, 18)]
    for n, h, w in geoms:
        xh, xw = math.ceil(h / 4), math.ceil(w / 4)
        ws, gs = make_hrfp_params(h + 3 * w)
        xp = make_feat(h * 7 + w, (n, 64, xh, xw))
        gen = torch.Generator(device="cuda").manual_seed(h + w)
        g1 = torch.randn((n, 64, xh, xw), device="cuda", generator=gen)
        g2 = torch.randn((n, 256, h // 2, w // 2), device="cuda", generator=gen)
        ref = _chain(xp, ws, gs, h, w, 0, g1, g2)
        got = _chain(xp, ws, gs, h, w, 3, g1, g2)
        for name, a, b, tol in zip(("OCout", "OCout_dec", "g_xp"), got, ref, (1e-2, 1e-2, 8e-2)):
            assert torch.isfinite(a).all(), (name, n, h, w)
            assert _l2(a, b) <= tol, (name, (n, h, w), _l2(a, b))
