"""NP+ (Normalization Perturbation Plus) on the sm_100a kernels.

Host-side mirror of `MRFPPlus.Normalization_Perturbation_Plus` (/root/reference/deepv3.py:268-277):
the two Gaussian draws are made here with torch's RNG in the reference's order (alpha first, then the
beta noise), everything else runs in one fused CUDA kernel per direction.
"""
import torch

from . import _lib


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


class _NPPlusFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, alpha, eps):
        lib = _lib.load()
        if not feat.is_cuda:
            raise _lib.MrfpError("NP+ kernels need a CUDA tensor (no CPU fallback)")
        if feat.dtype != torch.float32:
            raise _lib.MrfpError("NP+ kernels are fp32 (the reference runs fp32)")
        n, c, h, w = feat.shape
        x = feat.contiguous()
        alpha = alpha.reshape(n, c).to(torch.float32).contiguous()
        eps = eps.reshape(n, c).to(torch.float32).contiguous()
        out = torch.empty_like(x)
        mean = torch.empty((n, c), device=x.device, dtype=torch.float32)
        beta = torch.empty((n, c), device=x.device, dtype=torch.float32)
        ws_bytes = lib.mrfp_npplus_ws_bytes(n, c, h * w)
        ws = _lib.scratch(x.device, ws_bytes, "npplus")
        with torch.cuda.device(x.device):
            rc = lib.mrfp_npplus_fwd_f32(x.data_ptr(), alpha.data_ptr(), eps.data_ptr(), out.data_ptr(),
                                         mean.data_ptr(), beta.data_ptr(), ws.data_ptr(), ws_bytes,
                                         n, c, h * w, _stream_ptr(x))
        _lib.check(rc, "mrfp_npplus_fwd_f32")
        ctx.save_for_backward(alpha, eps, mean)
        ctx.beta = beta
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        alpha, eps, mean = ctx.saved_tensors
        n, c, h, w = gout.shape
        g = gout.contiguous()
        gin = torch.empty_like(g)
        ws_bytes = lib.mrfp_npplus_ws_bytes(n, c, h * w)
        ws = _lib.scratch(g.device, ws_bytes, "npplus")
        with torch.cuda.device(g.device):
            rc = lib.mrfp_npplus_bwd_f32(g.data_ptr(), alpha.data_ptr(), eps.data_ptr(), mean.data_ptr(),
                                         gin.data_ptr(), ws.data_ptr(), ws_bytes, n, c, h * w, _stream_ptr(g))
        _lib.check(rc, "mrfp_npplus_bwd_f32")
        return gin, None, None


class _ReluPsumFn(torch.autograd.Function):
    """y = relu(x) plus the plane sums of y (N*C doubles) taken in the same pass."""

    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        if not x.is_cuda or x.dtype != torch.float32:
            raise _lib.MrfpError("relu_with_plane_sums needs a CUDA fp32 tensor (no CPU fallback)")
        n, c, h, w = x.shape
        xc = x.contiguous()
        y = torch.empty_like(xc)
        psum = torch.empty((n, c), device=x.device, dtype=torch.float64)
        with torch.cuda.device(x.device):
            rc = lib.mrfp_relu_psum_f32(xc.data_ptr(), y.data_ptr(), psum.data_ptr(), n * c, h * w, _stream_ptr(xc))
        _lib.check(rc, "mrfp_relu_psum_f32")
        ctx.save_for_backward(y)
        ctx.mark_non_differentiable(psum)
        return y, psum

    @staticmethod
    def backward(ctx, gy, _gpsum):
        (y,) = ctx.saved_tensors
        return gy * (y > 0)


def relu_with_plane_sums(x: torch.Tensor):
    """ReLU that also returns sum_hw of its output per (n, c) plane — the statistics NP+ needs, taken by the producer."""
    return _ReluPsumFn.apply(x)


class _NPPlusPresummedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, psum, alpha, eps):
        lib = _lib.load()
        if not feat.is_cuda or feat.dtype != torch.float32:
            raise _lib.MrfpError("NP+ kernels need a CUDA fp32 tensor (no CPU fallback)")
        n, c, h, w = feat.shape
        x = feat.contiguous()
        alpha = alpha.reshape(n, c).to(torch.float32).contiguous()
        eps = eps.reshape(n, c).to(torch.float32).contiguous()
        psum = psum.reshape(n, c).to(torch.float64).contiguous()
        out = torch.empty_like(x)
        mean = torch.empty((n, c), device=x.device, dtype=torch.float32)
        ws_bytes = lib.mrfp_npplus_presummed_ws_bytes(n, c)
        ws = _lib.scratch(x.device, ws_bytes, "npplus_pre")
        with torch.cuda.device(x.device):
            rc = lib.mrfp_npplus_fwd_presummed_f32(x.data_ptr(), psum.data_ptr(), alpha.data_ptr(), eps.data_ptr(),
                                                   out.data_ptr(), mean.data_ptr(), None, ws.data_ptr(), ws_bytes,
                                                   n, c, h * w, _stream_ptr(x))
        _lib.check(rc, "mrfp_npplus_fwd_presummed_f32")
        ctx.save_for_backward(alpha, eps, mean)
        return out

    @staticmethod
    def backward(ctx, gout):
        gin, _, _ = _NPPlusFn.backward(ctx, gout)        # the ring kernel: needs the plane sums of gout
        return gin, None, None, None


def np_plus_presummed(feat: torch.Tensor, psum: torch.Tensor, alpha: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """NP+ (deepv3.py:268-277) as ONE streaming pass, given psum = sum_hw feat per plane from `relu_with_plane_sums`."""
    return _NPPlusPresummedFn.apply(feat, psum, alpha, eps)


def np_plus_with_draws(feat: torch.Tensor, alpha: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """NP+ with injected draws: alpha = torch.normal(1, .75), eps = torch.normal(0, .75), shape (N,C[,1,1])."""
    return _NPPlusFn.apply(feat, alpha, eps)


def draw_np_plus_factors(feat: torch.Tensor):
    """The reference's two draws, same order and shapes (deepv3.py:274-275), without its host syncs."""
    n, c = feat.shape[:2]
    alpha = torch.empty((n, c, 1, 1), device=feat.device, dtype=feat.dtype).normal_(1.0, 0.75)
    eps = torch.empty((n, c, 1, 1), device=feat.device, dtype=feat.dtype).normal_(0.0, 0.75)
    return alpha, eps


def normalization_perturbation_plus(feat: torch.Tensor) -> torch.Tensor:
    alpha, eps = draw_np_plus_factors(feat)
    return np_plus_with_draws(feat, alpha, eps)
