"""NP+ CUDA kernels (through the C ABI) vs the numpy oracle and the reference-generated fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import mrfp_oracle as O
from tests.common import GOLDEN, make_feat, make_draws

pytestmark = pytest.mark.gpu

# fp32 tolerance: |delta| <= 1e-5*max|ref| + 1e-5*|ref| forward, 1e-4 backward
FWD_TOL, BWD_TOL = 1e-5, 1e-4


def _run(feat, alpha, eps, gout=None):
    from mrfp_b200.npplus import np_plus_with_draws
    x = torch.from_numpy(feat).cuda().requires_grad_(gout is not None)
    y = np_plus_with_draws(x, torch.from_numpy(alpha).cuda(), torch.from_numpy(eps).cuda())
    gin = None
    if gout is not None:
        y.backward(torch.from_numpy(gout).cuda())
        gin = x.grad.cpu().numpy()
    return y.detach().cpu().numpy(), gin


def _close(got, ref, tol):
    ref = np.asarray(ref, dtype=np.float64)
    err = np.abs(got.astype(np.float64) - ref)
    bound = tol * np.abs(ref).max() + tol * np.abs(ref)
    assert (err <= bound).all(), f"max err {err.max():.3e} vs bound {bound.min():.3e}"


@pytest.mark.parametrize("i,name", list(enumerate("abcd")))
def test_vs_reference_fixture(i, name):
    g = np.load(os.path.join(GOLDEN, "npplus.npz"))
    shape = tuple(g[f"{name}_shape"])
    feat = make_feat(100 + i, shape)
    a, e = make_draws(200 + i, shape[0], shape[1])
    gout = np.random.default_rng(300 + i).standard_normal(shape).astype(np.float32)
    out, gin = _run(feat, a, e, gout)
    _close(out, g[f"{name}_out"], 2e-5)     # fixture itself is an fp32 computation
    _close(gin, g[f"{name}_gin"], 2e-4)


# incl. the stem / layer1 shapes of the other trunks (BASELINE configs 4-5, batch 2): R101 128ch@192^2, MobileNetV2
# 16ch@384^2 + 32ch@96^2, ShuffleNetV2 24ch@192^2 + 116ch@96^2
@pytest.mark.parametrize("shape", [(2, 64, 192, 192), (2, 256, 96, 96), (8, 16, 48, 48), (3, 24, 33, 31),
                                   (4, 116, 96, 96), (2, 3, 1, 1), (5, 7, 2, 2), (2, 600, 8, 8),
                                   (2, 128, 192, 192), (2, 16, 384, 384), (2, 32, 96, 96), (2, 24, 192, 192),
                                   (16, 64, 192, 192)])
def test_vs_oracle(shape):
    feat = make_feat(7, shape)
    a, e = make_draws(8, shape[0], shape[1])
    gout = np.random.default_rng(9).standard_normal(shape).astype(np.float32)
    out, gin = _run(feat, a, e, gout)
    f64 = feat.astype(np.float64)
    ref, mean, _ = O.np_plus_forward(f64, a.astype(np.float64), e.astype(np.float64))
    refg = O.np_plus_backward(gout.astype(np.float64), a.astype(np.float64), e.astype(np.float64), mean)
    if np.isnan(ref).any():
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        return
    _close(out, ref, FWD_TOL)
    _close(gin, refg, BWD_TOL)


def test_batch_of_one_is_nan_like_reference():
    feat = make_feat(1, (1, 4, 8, 8))
    a, e = make_draws(2, 1, 4)
    out, _ = _run(feat, a, e)
    assert np.isnan(out).all()


def test_equal_means_is_nan_like_reference():
    feat = np.ones((3, 4, 8, 8), dtype=np.float32)
    a, e = make_draws(3, 3, 4)
    out, _ = _run(feat, a, e)
    assert np.isnan(out).all()


def test_full_size_properties():
    """BASELINE config-2 shape (8,256,192,192): size-independent checks on the device."""
    from mrfp_b200.npplus import np_plus_with_draws
    torch.manual_seed(0)
    n, c, h, w = 8, 256, 192, 192
    x = torch.relu(torch.randn(n, c, h, w, device="cuda") * (0.5 + torch.rand(1, c, 1, 1, device="cuda"))
                   + torch.randn(1, c, 1, 1, device="cuda")).requires_grad_(True)
    alpha = 1 + 0.75 * torch.randn(n, c, 1, 1, device="cuda")
    eps = 0.75 * torch.randn(n, c, 1, 1, device="cuda")
    y = np_plus_with_draws(x, alpha, eps)
    # plane means: mean(out) = beta * mean(x); beta from the definition evaluated in fp64 on the device
    m = x.detach().double().mean((2, 3), keepdim=True)
    d = m.std(0, keepdim=True)
    beta = 1 + eps.double() * (d / d.max() * 1.5)
    ref = alpha.double() * x.detach().double() - alpha.double() * m + beta * m
    err = (y.double() - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item()
    g = torch.randn_like(x)
    y.backward(g)
    xr = x.detach().double().requires_grad_(True)
    mr = xr.mean((2, 3), keepdim=True)
    dr = mr.std(0, keepdim=True)
    yr = alpha.double() * xr - alpha.double() * mr + (1 + eps.double() * (dr / dr.max() * 1.5)) * mr
    yr.backward(g.double())
    errg = (x.grad.double() - xr.grad).abs().max().item()
    assert errg <= 1e-4 * xr.grad.abs().max().item()


@pytest.mark.parametrize("shape", [(8, 64, 96, 96), (4, 24, 33, 31), (8, 256, 192, 192)])
def test_workspace_protocol_and_determinism(shape):
    """Beyond its 64-byte control block the workspace may hold anything; the control block is zeroed once
    (mrfp_npplus_ws_init) and every launch leaves the phase-A queue counter zero (the phase-B counter is cleared at
    kernel entry), so one buffer serves launch after launch.  Zeros, 0xFF and
    random bytes give bit-identical outputs, forward and backward; a second pair of launches on the SAME buffer (no
    re-initialisation) does too."""
    from mrfp_b200 import _lib
    lib = _lib.load()
    n, c, h, w = shape
    torch.manual_seed(5)
    x = torch.relu(torch.randn(n, c, h, w, device="cuda"))
    g = torch.randn(n, c, h, w, device="cuda")
    alpha = 1 + 0.75 * torch.randn(n, c, device="cuda")
    eps = 0.75 * torch.randn(n, c, device="cuda")
    wsb = lib.mrfp_npplus_ws_bytes(n, c, h * w)
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for fill in ("zero", "ff", "rand", "rand"):
        ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
        if fill == "ff":
            ws.fill_(255)
        elif fill == "rand":
            ws.random_(0, 256)
        assert lib.mrfp_npplus_ws_init(ws.data_ptr(), wsb, st) == 0
        for rep in range(2):
            out = torch.empty_like(x); gin = torch.empty_like(x)
            mean = torch.empty(n, c, device="cuda"); beta = torch.empty(n, c, device="cuda")
            assert lib.mrfp_npplus_fwd_f32(x.data_ptr(), alpha.data_ptr(), eps.data_ptr(), out.data_ptr(), mean.data_ptr(),
                                           beta.data_ptr(), ws.data_ptr(), wsb, n, c, h * w, st) == 0
            assert lib.mrfp_npplus_bwd_f32(g.data_ptr(), alpha.data_ptr(), eps.data_ptr(), mean.data_ptr(), gin.data_ptr(),
                                           ws.data_ptr(), wsb, n, c, h * w, st) == 0
            torch.cuda.synchronize()
            assert int(ws[8:12].max()) == 0              # counter_a is zero again
            outs.append((out, gin, mean))
    for o, gi, m in outs[1:]:
        assert torch.equal(o, outs[0][0]) and torch.equal(gi, outs[0][1]) and torch.equal(m, outs[0][2])
    assert lib.mrfp_npplus_ws_init(None, wsb, st) == -1
    assert lib.mrfp_npplus_ws_init(ws.data_ptr(), 8, st) == -3


@pytest.mark.parametrize("shape", [(4, 64, 96, 96), (3, 24, 33, 31)])
def test_cuda_graph_replay_of_forward_and_backward(shape):
    """NP+ forward + backward captured into a CUDA graph and replayed on new inputs equals the eager calls (the ring
    kernel's queue counters are restored by the kernel itself, nothing per-launch comes from the host)."""
    from mrfp_b200.npplus import np_plus_with_draws
    n, c, h, w = shape
    torch.manual_seed(11)
    xs = torch.relu(torch.randn(n, c, h, w, device="cuda"))
    gs = torch.randn(n, c, h, w, device="cuda")
    a = 1 + 0.75 * torch.randn(n, c, device="cuda")
    e = 0.75 * torch.randn(n, c, device="cuda")

    def run(x_, g_):
        xr = x_.detach().requires_grad_(True)
        y = np_plus_with_draws(xr, a, e)
        (gx,) = torch.autograd.grad(y, xr, g_)
        return y, gx

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):                               # warm-up on the capture stream: scratch buffers, attributes
            run(xs, gs)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        y_g, gx_g = run(xs, gs)
    for it in range(3):
        xs.copy_(torch.relu(torch.randn(n, c, h, w, device="cuda")))
        gs.copy_(torch.randn(n, c, h, w, device="cuda"))
        graph.replay()
        torch.cuda.synchronize()
        y_e, gx_e = run(xs, gs)
        assert torch.equal(y_g, y_e) and torch.equal(gx_g, gx_e), it


@pytest.mark.parametrize("shape", [(2, 256, 192, 192), (3, 24, 33, 31), (4, 116, 96, 96), (8, 64, 48, 48)])
def test_producer_fused_path_equals_the_ring_kernel(shape):
    """relu_with_plane_sums + np_plus_presummed (NP+ statistics taken by the producer's ReLU) vs
    torch.relu + the standalone NP+ kernel: same output and the same gradient into the pre-ReLU tensor."""
    from mrfp_b200.npplus import np_plus_with_draws, np_plus_presummed, relu_with_plane_sums
    n, c, h, w = shape
    rng = np.random.default_rng(41)
    pre = torch.from_numpy(rng.standard_normal(shape).astype(np.float32) + 0.3).cuda()
    a, e = make_draws(42, n, c)
    a, e = torch.from_numpy(a).cuda(), torch.from_numpy(e).cuda()
    g = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).cuda()

    x1 = pre.clone().requires_grad_(True)
    y1, psum = relu_with_plane_sums(x1)
    o1 = np_plus_presummed(y1, psum, a, e)
    o1.backward(g)
    x2 = pre.clone().requires_grad_(True)
    y2 = torch.relu(x2)
    o2 = np_plus_with_draws(y2, a, e)
    o2.backward(g)

    assert torch.equal(y1, y2)
    ref_sum = y2.double().sum((2, 3))
    assert torch.allclose(psum, ref_sum, rtol=1e-6, atol=1e-6 * ref_sum.abs().max().item())
    _close(o1.detach().cpu().numpy(), o2.detach().cpu().numpy(), 2e-6)
    _close(x1.grad.cpu().numpy(), x2.grad.cpu().numpy(), 2e-6)
    # and against the oracle
    ref, _, _ = O.np_plus_forward(y2.detach().cpu().numpy().astype(np.float64), a.cpu().numpy().astype(np.float64),
                                  e.cpu().numpy().astype(np.float64))
    _close(o1.detach().cpu().numpy(), ref, FWD_TOL)


def test_rng_stream_matches_reference_draw_order():
    """Same generator state -> same alpha/eps as the reference's two torch.normal calls (deepv3.py:274-275)."""
    from mrfp_b200.npplus import draw_np_plus_factors
    feat = torch.zeros(4, 64, 2, 2, device="cuda")
    torch.manual_seed(123)
    a, e = draw_np_plus_factors(feat)
    torch.manual_seed(123)
    ones = torch.ones(4, 64, 1, 1, device="cuda")
    a_ref = torch.normal(ones, 0.75 * ones)
    e_ref = torch.normal(torch.zeros_like(ones), 0.75 * ones)
    assert torch.allclose(a, a_ref, atol=1e-6) and torch.allclose(e, e_ref, atol=1e-6)


def test_bad_arguments_return_codes():
    from mrfp_b200 import _lib
    lib = _lib.load()
    assert lib.mrfp_npplus_fwd_f32(0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 4, 0) == -1
    x = torch.zeros(16, device="cuda")
    p = x.data_ptr()
    assert lib.mrfp_npplus_fwd_f32(p, p, p, p, p, p, p, 0, 2, 2, 4, 0) == -3
    assert lib.mrfp_npplus_fwd_f32(p, p, p, p, p, p, p, 1 << 20, 0, 2, 4, 0) == -2
    assert b"workspace" in lib.mrfp_strerror(-3)
