"""Single passes of the HRFP chain, each pinned on its own at production row widths (768^2 geometry).

The chain tests cannot separate a small systematic error of one element-wise kernel from the bf16 storage noise of
eight chained stages, so every product kernel of `csrc/bn_ring.cu` (bulk-copy forward BN/ReLU/resample, BN-backward
reduce, BN-backward apply), the identity-stream apply kernel, their LDG counterparts in `csrc/hrfp.cu` and the
`stmatrix` layout kernel run here ALONE through the library's test hooks, on identical bf16 (or fp32) inputs, against
an fp64 evaluation of the same formula (deepv3.py:320-327: F.interpolate(nearest) -> BatchNorm2d(train) -> ReLU and
its autograd backward).  Tolerance: one ulp of the storage type on the result plus fp32 rounding of the terms.
"""
import numpy as np
import pytest
import torch

from oracle import mrfp_oracle as O

pytestmark = pytest.mark.gpu

H = W = 768
XH = XW = 192
BF16_ULP = 2.0 ** -8          # relative spacing of bf16 (8 significand bits): |round(x) - x| <= ulp / 2
FP32_EPS = 2.0 ** -23


def _plan(n, mode):
    from mrfp_b200.hrfp import get_plan
    return get_plan(n, 64, XH, XW, H, W, torch.device("cuda"), mode)


def _geom():
    return O.hrfp_geometry(H, W, XH, XW)


def _stage_inputs(k, n, dtype, seed):
    """Random tensors of stage k: conv output y (N,ch,cw,C), upstream gradient dA (N,oh,ow,C), BN parameters."""
    st = _geom()[k]
    c = O.HRFP_LAYERS[k][1]
    g = torch.Generator(device="cuda").manual_seed(seed)
    y = (torch.randn(n, st.conv_h, st.conv_w, c, device="cuda", generator=g) * 1.5 + 0.3).to(dtype)
    oh, ow = len(st.idx_h), len(st.idx_w)
    dA = torch.randn(n, oh, ow, c, device="cuda", generator=g).to(dtype)
    gamma = (0.5 * torch.randn(c, device="cuda", generator=g)).float()
    beta = (0.1 * torch.randn(c, device="cuda", generator=g)).float()
    # batch statistics of the RESAMPLED tensor, as the conv epilogue leaves them (fp64 -> fp32 table)
    idx_h = torch.from_numpy(np.asarray(st.idx_h)).cuda()
    idx_w = torch.from_numpy(np.asarray(st.idx_w)).cuda()
    r = y.double()[:, idx_h][:, :, idx_w]
    mean = r.mean((0, 1, 2)); var = r.var((0, 1, 2), unbiased=False)
    invstd = 1.0 / torch.sqrt(var + 1e-5)
    scale = (gamma.double() * invstd).float()
    shift = (beta.double() - mean * scale.double()).float()
    stats = torch.zeros(4, 256, device="cuda")
    stats[0, :c] = mean.float(); stats[1, :c] = invstd.float(); stats[2, :c] = scale; stats[3, :c] = shift
    return st, c, y, dA, gamma, stats, idx_h, idx_w, oh, ow


def _call(plan, k, op, variant, y, in2, out, stats, gamma, acc):
    from mrfp_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.mrfp_debug_stage_op(plan.handle, plan.lut.data_ptr(), k, op, variant, y.data_ptr(),
                                 None if in2 is None else in2.data_ptr(), None if out is None else out.data_ptr(),
                                 stats.data_ptr(), None if gamma is None else gamma.data_ptr(),
                                 None if acc is None else acc.data_ptr(), st)
    _lib.check(rc, "mrfp_debug_stage_op")
    torch.cuda.synchronize()


CASES = [(2, "up x1.2, 128 ch, 277->332"), (3, "up, 256 ch, 332->384"), (4, "identity, 128 ch, 384"),
         (5, "down x0.838, 64 ch, 384->321"), (7, "down, 64 ch, 256->192")]


@pytest.mark.parametrize("mode", [2, 1])
@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("k", [c[0] for c in CASES])
def test_forward_bn_relu_resample(k, variant, mode):
    """A_next = ReLU(scale * Y[idx_h][idx_w] + shift), stored in the plan's element type."""
    n = 2
    dtype = torch.bfloat16 if mode == 2 else torch.float32
    plan = _plan(n, mode)
    st, c, y, _, _, stats, idx_h, idx_w, oh, ow = _stage_inputs(k, n, dtype, 100 + k)
    out = torch.empty(n, oh, ow, c, device="cuda", dtype=dtype)
    _call(plan, k, 0, variant, y, None, out, stats, None, None)
    ref = torch.relu(stats[2, :c].double() * y.double()[:, idx_h][:, :, idx_w] + stats[3, :c].double())
    err = (out.double() - ref).abs()
    ulp = BF16_ULP if mode == 2 else 2 * FP32_EPS
    bound = ulp * ref.abs() + 4 * FP32_EPS * (stats[2, :c].double().abs() * y.double()[:, idx_h][:, :, idx_w].abs() + stats[3, :c].double().abs())
    assert bool((err <= bound).all()), float((err - bound).max())


@pytest.mark.parametrize("mode", [2, 1])
@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("k", [c[0] for c in CASES])
def test_bn_backward_reduce(k, variant, mode):
    """U1[c] = sum mask * dA, U2[c] = sum mask * dA * y over the resampled tensor, mask = [scale * y + shift > 0]."""
    n = 2
    dtype = torch.bfloat16 if mode == 2 else torch.float32
    plan = _plan(n, mode)
    st, c, y, dA, _, stats, idx_h, idx_w, oh, ow = _stage_inputs(k, n, dtype, 200 + k)
    acc = torch.zeros(2, 256, device="cuda", dtype=torch.float64)
    _call(plan, k, 1, variant, y, dA, None, stats, None, acc)
    yr = y.double()[:, idx_h][:, :, idx_w]
    # the sign of an fp32 fma equals the sign of the exact value, so the fp64 mask is the kernel's mask
    t = torch.where(stats[2, :c].double() * yr + stats[3, :c].double() > 0, dA.double(), torch.zeros((), device="cuda", dtype=torch.float64))
    u1, u2 = t.sum((0, 1, 2)), (t * yr).sum((0, 1, 2))
    a1, a2 = t.abs().sum((0, 1, 2)), (t * yr).abs().sum((0, 1, 2))
    # fp32 partial sums of a few hundred terms per thread, double across threads
    assert bool(((acc[0, :c] - u1).abs() <= 2e-5 * a1 + 1e-9).all()), float(((acc[0, :c] - u1).abs() / a1).max())
    assert bool(((acc[1, :c] - u2).abs() <= 2e-5 * a2 + 1e-9).all()), float(((acc[1, :c] - u2).abs() / a2).max())
    assert c == 256 or float(acc[:, c:].abs().max()) == 0.0


@pytest.mark.parametrize("mode", [2, 1])
@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("k", [c[0] for c in CASES])
def test_bn_backward_apply(k, variant, mode):
    """dY[src] = invstd * (gamma * mask * sum_replicas dA - cnt * (M1 + xhat * M2))  — BatchNorm2d(train) backward composed
    with the adjoint of the nearest resample (sum over the replicas of a source pixel; 0 replicas -> only the mean terms
    vanish too because cnt = 0)."""
    n = 2
    dtype = torch.bfloat16 if mode == 2 else torch.float32
    plan = _plan(n, mode)
    st, c, y, dA, gamma, stats, idx_h, idx_w, oh, ow = _stage_inputs(k, n, dtype, 300 + k)
    yr = y.double()[:, idx_h][:, :, idx_w]
    sc, sf = stats[2, :c].double(), stats[3, :c].double()
    mean, invstd = stats[0, :c].double(), stats[1, :c].double()
    mask = sc * yr + sf > 0
    t = torch.where(mask, dA.double(), torch.zeros((), device="cuda", dtype=torch.float64))
    acc = torch.zeros(2, 256, device="cuda", dtype=torch.float64)
    acc[0, :c] = t.sum((0, 1, 2)); acc[1, :c] = (t * yr).sum((0, 1, 2))
    out = torch.full((n, st.conv_h, st.conv_w, c), float("nan"), device="cuda", dtype=dtype)
    _call(plan, k, 2, variant, y, dA, out, stats, gamma, acc)
    # fp64 reference in the reference's own terms: BN backward on the resampled tensor, then the resample adjoint
    count = float(n * oh * ow)
    g64 = gamma.double()
    dxh = t * g64
    m1 = dxh.sum((0, 1, 2)) / count
    m2 = (dxh * (yr - mean) * invstd).sum((0, 1, 2)) / count
    dr = invstd * (dxh - m1 - (yr - mean) * invstd * m2)                      # (N, oh, ow, C)
    tmp = torch.zeros(n, st.conv_h, ow, c, device="cuda", dtype=torch.float64).index_add_(1, idx_h, dr)
    ref = torch.zeros(n, st.conv_h, st.conv_w, c, device="cuda", dtype=torch.float64).index_add_(2, idx_w, tmp)
    # magnitude of the terms the kernel combines in fp32: P*mask*SdA and cnt*(R*y + Q)
    cnt_h = torch.bincount(idx_h, minlength=st.conv_h).double()
    cnt_w = torch.bincount(idx_w, minlength=st.conv_w).double()
    cnt = cnt_h[:, None] * cnt_w[None, :]
    sda = torch.zeros_like(ref).index_add_(2, idx_w, torch.zeros(n, st.conv_h, ow, c, device="cuda", dtype=torch.float64).index_add_(1, idx_h, t.abs()))
    r_ = invstd * invstd * m2
    q_ = invstd * m1 - mean * r_
    mag = (invstd * g64).abs() * sda + cnt[None, :, :, None] * ((r_ * y.double()).abs() + q_.abs())
    err = (out.double() - ref).abs()
    ulp = BF16_ULP if mode == 2 else 2 * FP32_EPS
    bound = ulp * ref.abs() + 8 * FP32_EPS * mag + 1e-12
    assert bool(torch.isfinite(out.float()).all())
    assert bool((err <= bound).all()), float((err - bound).max())


@pytest.mark.parametrize("shape", [(2, 64, 192 * 192), (2, 256, 384 * 384), (3, 128, 40 * 28)])
def test_nchw_to_nhwc_stmatrix_kernel_is_exact(shape):
    """fp32 NCHW -> bf16 NHWC through `stmatrix.trans`: every element is the round-to-nearest-even bf16 of its source
    (bit-exact against torch's conversion), the plane totals of the folded NP+ agree with an fp64 sum; the generic
    fp32-tile kernel gives the same bits."""
    from mrfp_b200 import _lib
    lib = _lib.load()
    n, c, hw = shape
    torch.manual_seed(7)
    src = torch.randn(n, c, hw, device="cuda") * 3
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for variant in (1, 0):
        dst = torch.empty(n, hw, c, device="cuda", dtype=torch.bfloat16)
        psum = torch.zeros(n, c, device="cuda", dtype=torch.float64)
        _lib.check(lib.mrfp_debug_nchw_to_nhwc(src.data_ptr(), dst.data_ptr(), n, c, c, hw, 2, variant, psum.data_ptr(), st), "layout")
        torch.cuda.synchronize()
        assert torch.equal(dst, src.permute(0, 2, 1).to(torch.bfloat16))
        ref = src.double().sum(2)
        assert bool(((psum - ref).abs() <= 1e-6 * src.double().abs().sum(2)).all())
        outs.append(dst)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("esize", [2, 4])
def test_nchw_to_nhwc_pads_a_narrow_stem_with_zero_channels(esize):
    """24 source channels stored as 64 (ShuffleNetV2 stem): real channels converted, padding zero."""
    from mrfp_b200 import _lib
    lib = _lib.load()
    n, c, cd, hw = 2, 24, 64, 37 * 29
    torch.manual_seed(8)
    src = torch.randn(n, c, hw, device="cuda")
    dt = torch.bfloat16 if esize == 2 else torch.float32
    dst = torch.full((n, hw, cd), float("nan"), device="cuda", dtype=dt)
    psum = torch.zeros(n, c, device="cuda", dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.mrfp_debug_nchw_to_nhwc(src.data_ptr(), dst.data_ptr(), n, c, cd, hw, esize, 1, psum.data_ptr(), st), "layout")
    torch.cuda.synchronize()
    assert torch.equal(dst[:, :, :c], src.permute(0, 2, 1).to(dt))
    assert float(dst[:, :, c:].abs().max()) == 0.0
    assert bool(((psum - src.double().sum(2)).abs() <= 1e-6 * src.double().abs().sum(2)).all())
