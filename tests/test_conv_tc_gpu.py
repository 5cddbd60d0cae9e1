"""tcgen05 implicit-GEMM conv kernel alone (debug hook of the C library) vs torch conv2d on the same bf16 inputs."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _conv_tc(x_nchw, w_oihw, dil, cnt_h=None, cnt_w=None):
    from mrfp_b200 import _lib
    lib = _lib.load()
    fn = lib.mrfp_debug_conv3x3_bf16
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 4
    n, cin, h, w = x_nchw.shape
    cout = w_oihw.shape[0]
    x = x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    wp = w_oihw.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous().to(torch.bfloat16)   # [tap][co][ci]
    out = torch.full((n, h, w, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    acc = torch.zeros(2, 256, device="cuda", dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    rc = fn(x.data_ptr(), wp.data_ptr(), out.data_ptr(), n, h, w, cin, cout, dil,
            None if cnt_h is None else cnt_h.data_ptr(), None if cnt_w is None else cnt_w.data_ptr(),
            acc.data_ptr() if cnt_h is not None else None, st)
    assert rc == 0, rc
    torch.cuda.synchronize()
    return out.permute(0, 3, 1, 2).float(), acc, x.permute(0, 3, 1, 2).float(), wp.reshape(3, 3, cout, cin).permute(2, 3, 0, 1).float()


@pytest.mark.parametrize("cin,cout,dil", [(64, 64, 1), (64, 64, 2), (64, 128, 2), (128, 256, 2), (256, 128, 1), (128, 64, 1)])
@pytest.mark.parametrize("shape", [(2, 8, 16), (2, 20, 37), (1, 50, 61)])
def test_conv_matches_torch(cin, cout, dil, shape):
    n, h, w = shape
    torch.manual_seed(cin + cout + dil + h)
    x = torch.relu(torch.randn(n, cin, h, w, device="cuda"))
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * (2.0 / (9 * cin)) ** 0.5
    cnt_h = torch.zeros(((h + 15) // 16 + 1) * 16, dtype=torch.int32, device="cuda")
    cnt_w = torch.zeros(((w + 15) // 16 + 1) * 16, dtype=torch.int32, device="cuda")
    cnt_h[:h] = torch.randint(0, 3, (h,), device="cuda", dtype=torch.int32)
    cnt_w[:w] = torch.randint(0, 3, (w,), device="cuda", dtype=torch.int32)
    got, acc, xb, wb = _conv_tc(x, wt, dil, cnt_h, cnt_w)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(xb.double(), wb.double(), padding=dil, dilation=dil)
    assert not torch.isnan(got).any()
    err = (got.double() - ref).abs().max().item()
    # bf16 output rounding: half an ulp = 2^-9 relative
    assert err <= 2 ** -8 * ref.abs().max().item() + 1e-6, err
    # the BN statistics are taken over the STORED (bf16) conv output, weighted by the replication counts
    wgt = (cnt_h[:h].double()[:, None] * cnt_w[:w].double()[None, :])[None, None]
    gd = got.double()
    s1 = (gd * wgt).sum((0, 2, 3)); s2 = (gd * gd * wgt).sum((0, 2, 3))
    assert torch.allclose(acc[0, :cout], s1, rtol=1e-4, atol=1e-4 * s1.abs().max().item())
    assert torch.allclose(acc[1, :cout], s2, rtol=1e-4, atol=1e-4 * s2.abs().max().item())


def test_conv_without_stats_full_size_linearity():
    """D1 shape of the reference chain (256->128 @384^2), batch 2: conv(a*x) = a*conv(x) and agreement with cuDNN."""
    torch.manual_seed(0)
    x = torch.relu(torch.randn(2, 256, 384, 384, device="cuda"))
    wt = torch.randn(128, 256, 3, 3, device="cuda") * (2.0 / (9 * 256)) ** 0.5
    got, _, xb, wb = _conv_tc(x, wt, 1)
    got2, _, _, _ = _conv_tc(2 * x, wt, 1)
    assert torch.equal(got2, 2 * got)                 # exact: power-of-two scaling commutes with bf16 rounding
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(xb, wb, padding=1)
    assert (got - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item() + 1e-5


def _nearest_idx(src, dst):
    scale = torch.tensor(src / dst, dtype=torch.float32)
    return torch.clamp(torch.floor(torch.arange(dst, dtype=torch.float32) * scale).to(torch.int64), max=src - 1).cuda()


@pytest.mark.parametrize("cin,cout,dil", [(64, 64, 1), (64, 64, 2), (64, 128, 2), (128, 256, 2), (256, 128, 1), (128, 64, 1)])
@pytest.mark.parametrize("geom", [(2, 16, 16, 16, 16), (2, 20, 37, 17, 31), (1, 50, 61, 60, 75), (2, 33, 24, 28, 20)])
def test_gathered_conv_matches_torch(cin, cout, dil, geom):
    """conv3x3_gather_kernel alone (csrc/conv_gather.cu): Conv2d(ReLU(scale * nearest(Y_prev) + shift)) with the operand
    built on chip, against torch on the same bf16 operand (rounded once, like the materialised A_k of the unfused path):
    up-sampling, down-sampling and identity resamples, ragged tiles, and the BN statistics of the stored output."""
    import ctypes as C
    from mrfp_b200 import _lib
    n, h, w, sh, sw = geom
    torch.manual_seed(cin + cout + dil + h + sh)
    fn = _lib.load().mrfp_debug_conv3x3_gather_fwd
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int] * 6 + [C.c_void_p] * 4
    y = torch.randn(n, sh, sw, cin, device="cuda").to(torch.bfloat16)
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * (2.0 / (9 * cin)) ** 0.5
    wp = wt.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous().to(torch.bfloat16)
    stats = torch.zeros(4, 256, device="cuda")
    stats[2, :cin] = 0.5 + torch.rand(cin, device="cuda"); stats[3, :cin] = 0.3 * torch.randn(cin, device="cuda")
    ih, iw = _nearest_idx(sh, h), _nearest_idx(sw, w)
    ih32, iw32 = ih.to(torch.int32).contiguous(), iw.to(torch.int32).contiguous()
    cnt_h = torch.zeros(((h + 15) // 16 + 1) * 16, dtype=torch.int32, device="cuda")
    cnt_w = torch.zeros(((w + 15) // 16 + 1) * 16, dtype=torch.int32, device="cuda")
    cnt_h[:h] = torch.randint(0, 3, (h,), device="cuda", dtype=torch.int32)
    cnt_w[:w] = torch.randint(0, 3, (w,), device="cuda", dtype=torch.int32)
    acc = torch.zeros(2, 256, device="cuda", dtype=torch.float64)
    out = torch.full((n, h, w, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = fn(y.data_ptr(), sh, sw, ih32.data_ptr(), iw32.data_ptr(), stats.data_ptr(), wp.data_ptr(), out.data_ptr(), n, h, w, cin,
            cout, dil, cnt_h.data_ptr(), cnt_w.data_ptr(), acc.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    a = torch.relu(torch.addcmul(stats[3, :cin], y.float().index_select(1, ih).index_select(2, iw), stats[2, :cin]))
    a = a.to(torch.bfloat16).permute(0, 3, 1, 2).double()                    # fmaf(scale, y, shift) -> one bf16 rounding
    wb = wp.reshape(3, 3, cout, cin).permute(2, 3, 0, 1).double()
    ref = F.conv2d(a, wb, padding=dil, dilation=dil)
    got = out.permute(0, 3, 1, 2).double()
    assert not torch.isnan(got).any()
    # bf16 output rounding (2^-9 relative) + the rare operand that rounds the other way (fma vs mul-add of the reference)
    assert (got - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item() + 1e-6
    assert float((got - ref).norm() / ref.norm()) <= 3e-3
    wgt = (cnt_h[:h].double()[:, None] * cnt_w[:w].double()[None, :])[None, None]
    s1 = (got * wgt).sum((0, 2, 3)); s2 = (got * got * wgt).sum((0, 2, 3))
    assert torch.allclose(acc[0, :cout], s1, rtol=1e-4, atol=1e-4 * s1.abs().max().item())
    assert torch.allclose(acc[1, :cout], s2, rtol=1e-4, atol=1e-4 * s2.abs().max().item())


@pytest.mark.parametrize("cin,cout,dil,geom,add", [
    (128, 256, 1, (2, 40, 56, 40, 56), True),      # stage-4 dgrad: identity resample, OCout_dec gradient added in the epilogue
    (64, 128, 1, (2, 33, 24, 28, 20), False),      # down-sampling: some source pixels have no replica
    (64, 64, 2, (2, 50, 61, 40, 49), False),
    (64, 64, 1, (2, 40, 56, 48, 67), False),       # up-sampling x1.2: up to 2 x 2 replicas per source pixel (extras stage)
    (64, 64, 1, (1, 50, 61, 60, 73), False),
])
def test_gathered_dgrad_matches_torch(cin, cout, dil, geom, add):
    """conv3x3_gather_kernel<bwd | bwd-rep> alone: Conv2d(dY) with dY = BN-backward apply of (dA, Y) under the nearest
    adjoint built on chip, against torch in fp64 on the same bf16 operands (dY rounded to bf16 once, as the materialised
    dY of the unfused path)."""
    import ctypes as C
    from mrfp_b200 import _lib
    n, h, w, oh, ow = geom                          # (h, w): conv resolution of the stage; (oh, ow): its resampled output
    torch.manual_seed(cin + cout + dil + h + oh)
    fn = _lib.load().mrfp_debug_conv3x3_gather_bwd
    fn.restype = C.c_int
    fn.argtypes = ([C.c_void_p] * 2 + [C.c_int] * 2 + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3 + [C.c_double] +
                   [C.c_void_p] * 2 + [C.c_int] * 6 + [C.c_void_p] * 2)
    y = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    dA = torch.randn(n, oh, ow, cin, device="cuda").to(torch.bfloat16)
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * (2.0 / (9 * cin)) ** 0.5
    wp = wt.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous().to(torch.bfloat16)
    stats = torch.zeros(4, 256, device="cuda")
    stats[0, :cin] = 0.1 * torch.randn(cin, device="cuda"); stats[1, :cin] = 0.5 + torch.rand(cin, device="cuda")
    gamma = (0.5 * torch.randn(cin, device="cuda")).contiguous()
    stats[2, :cin] = gamma * stats[1, :cin]; stats[3, :cin] = -stats[0, :cin] * stats[2, :cin]
    count = float(n * oh * ow)
    acc = torch.zeros(2, 256, device="cuda", dtype=torch.float64)
    acc[0, :cin] = count * 0.05 * torch.randn(cin, device="cuda", dtype=torch.float64)
    acc[1, :cin] = count * 0.05 * torch.randn(cin, device="cuda", dtype=torch.float64)
    ih, iw = _nearest_idx(h, oh).cpu(), _nearest_idx(w, ow).cpu()      # dst -> src of the forward resample
    lo_h = torch.searchsorted(ih, torch.arange(h + 1)).to(torch.int32).contiguous()
    lo_w = torch.searchsorted(iw, torch.arange(w + 1)).to(torch.int32).contiguous()
    max_rep = int(max((lo_h[1:] - lo_h[:-1]).max(), (lo_w[1:] - lo_w[:-1]).max()))
    assert max_rep <= 2
    lo_hd, lo_wd = lo_h.cuda(), lo_w.cuda()
    addt = torch.randn(n, h, w, cout, device="cuda").to(torch.bfloat16) if add else None
    out = torch.full((n, h, w, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = fn(y.data_ptr(), dA.data_ptr(), oh, ow, lo_hd.data_ptr(), lo_wd.data_ptr(), lo_h.data_ptr(), lo_w.data_ptr(), max_rep,
            stats.data_ptr(), gamma.data_ptr(), acc.data_ptr(), count, wp.data_ptr(), out.data_ptr(), n, h, w, cin, cout, dil,
            None if addt is None else addt.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    # reference: the adjoint of the nearest resample sums the replicas of each source pixel; BN backward with the given sums
    S = torch.zeros(n, h, ow, cin, device="cuda", dtype=torch.float64).index_add_(1, ih.cuda(), dA.double())
    S = torch.zeros(n, h, w, cin, device="cuda", dtype=torch.float64).index_add_(2, iw.cuda(), S)
    cnt = ((lo_hd[1:] - lo_hd[:-1]).double()[:, None] * (lo_wd[1:] - lo_wd[:-1]).double()[None, :])[None, :, :, None]
    mean, invstd, gm = stats[0, :cin].double(), stats[1, :cin].double(), gamma.double()
    S1, S2 = acc[0, :cin], invstd * (acc[1, :cin] - mean * acc[0, :cin])
    r = invstd * invstd * (gm * S2 / count)
    P, Q, R = (invstd * gm).float(), (invstd * (gm * S1 / count) - mean * r).float(), r.float()
    yf = y.float()
    mask = torch.addcmul(stats[3, :cin], yf, stats[2, :cin]) > 0
    dY = P.double() * torch.where(mask, S, torch.zeros_like(S)) - cnt * (R.double() * yf.double() + Q.double())
    dY = dY.to(torch.bfloat16).permute(0, 3, 1, 2).double()
    wb = wp.reshape(3, 3, cout, cin).permute(2, 3, 0, 1).double()
    ref = F.conv2d(dY, wb, padding=dil, dilation=dil)
    if addt is not None:
        ref = ref + addt.permute(0, 3, 1, 2).double()
    got = out.permute(0, 3, 1, 2).double()
    assert not torch.isnan(got).any()
    assert (got - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item() + 1e-6
    assert float((got - ref).norm() / ref.norm()) <= 3e-3
