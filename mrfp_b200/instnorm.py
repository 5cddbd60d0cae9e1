"""InstanceNorm2d(affine) + ReLU of the trunk next to the MRFP insertion points, on the sm_100a cluster kernels.

Host-side mirror of `nn.InstanceNorm2d(C, affine=True)` followed by `nn.ReLU` as the reference's trunk applies
them with wt_layer=[0,0,4,4,4,0,0] (/root/reference/network/Resnet.py:534-536 + :596-598 stem;
:176-178 + :218-225 last Bottleneck of layer1 and layer2) — SURVEY.md §8f-3.  `nn.InstanceNorm2d` keeps no running
statistics by default, so training and eval compute the same thing.  The layer1 instance is the producer of NP+
call 2 (deepv3.py:334-335): `want_plane_sums=True` makes the same pass leave sum_hw of its output per plane.
"""
import torch

from . import _lib


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


class _InstNormReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, relu, want_plane_sums):
        lib = _lib.load()
        if not x.is_cuda or x.dtype != torch.float32:
            raise _lib.MrfpError("instance_norm_relu needs a CUDA fp32 tensor (no CPU fallback)")
        n, c, h, w = x.shape
        xc = x.contiguous()
        wt = None if weight is None else weight.detach().to(torch.float32).contiguous()
        bs = None if bias is None else bias.detach().to(torch.float32).contiguous()
        y = torch.empty_like(xc)
        mean = torch.empty((n, c), device=x.device, dtype=torch.float32)
        invstd = torch.empty((n, c), device=x.device, dtype=torch.float32)
        psum = torch.empty((n, c), device=x.device, dtype=torch.float64) if want_plane_sums else None
        with torch.cuda.device(x.device):
            rc = lib.mrfp_instnorm_fwd_f32(xc.data_ptr(), None if wt is None else wt.data_ptr(),
                                           None if bs is None else bs.data_ptr(), y.data_ptr(), mean.data_ptr(),
                                           invstd.data_ptr(), None if psum is None else psum.data_ptr(),
                                           n, c, h * w, float(eps), int(bool(relu)), _stream_ptr(xc))
        _lib.check(rc, "mrfp_instnorm_fwd_f32")
        ctx.save_for_backward(xc, wt, bs, mean, invstd)
        ctx.relu = bool(relu)
        ctx.has_affine = (weight is not None, bias is not None)
        if psum is None:
            psum = torch.empty(0, device=x.device, dtype=torch.float64)
        ctx.mark_non_differentiable(psum)
        return y, psum

    @staticmethod
    def backward(ctx, gy, _gpsum):
        lib = _lib.load()
        x, wt, bs, mean, invstd = ctx.saved_tensors
        n, c, h, w = x.shape
        g = gy.contiguous()
        gx = torch.empty_like(x)
        dg = torch.empty((n, c), device=x.device, dtype=torch.float32)
        db = torch.empty((n, c), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            rc = lib.mrfp_instnorm_bwd_f32(g.data_ptr(), x.data_ptr(), None if wt is None else wt.data_ptr(),
                                           None if bs is None else bs.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                           gx.data_ptr(), dg.data_ptr(), db.data_ptr(), n, c, h * w, int(ctx.relu),
                                           _stream_ptr(x))
        _lib.check(rc, "mrfp_instnorm_bwd_f32")
        gw = dg.sum(0) if ctx.has_affine[0] and ctx.needs_input_grad[1] else None
        gb = db.sum(0) if ctx.has_affine[1] and ctx.needs_input_grad[2] else None
        return gx, gw, gb, None, None, None


def instance_norm_relu(x, weight=None, bias=None, eps=1e-5, relu=True, want_plane_sums=False):
    """relu(F.instance_norm(x, weight=weight, bias=bias, eps=eps)) in one on-chip pass per plane.
    Returns y, or (y, psum) with psum[n, c] = sum_hw y (float64) when want_plane_sums."""
    y, psum = _InstNormReluFn.apply(x, weight, bias, eps, relu, want_plane_sums)
    return (y, psum) if want_plane_sums else y


def module_instance_norm_relu(module: torch.nn.InstanceNorm2d, x, relu=True, want_plane_sums=False):
    """The fused op with the parameters of an `nn.InstanceNorm2d` module (state_dict keys unchanged).  Modules that
    track running statistics are not what the reference builds (Resnet.py:176-178, :534-536) and are refused."""
    if module.track_running_stats:
        raise _lib.MrfpError("instance_norm_relu: track_running_stats=True is not supported")
    return instance_norm_relu(x, module.weight, module.bias, module.eps, relu, want_plane_sums)


class _InstNormReluNpFn(torch.autograd.Function):
    """out = NP+(ReLU(InstanceNorm(x))) with injected NP+ draws (NP+ call 2, deepv3.py:334-335, and its
    producer, Resnet.py:218-225, as one autograd node).  Forward: the IN + ReLU kernel also leaves the plane sums of its
    output, NP+ is one streaming pass over it.  Backward: the plane totals of the incoming gradient, the NP+ backward
    coefficients, and the InstanceNorm backward that applies them on load — the gradient of the NP+ input is never
    materialised, and the normalised tensor is not kept for backward (the ReLU mask is recomputed from x)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, alpha, eps_draw):
        lib = _lib.load()
        if not x.is_cuda or x.dtype != torch.float32:
            raise _lib.MrfpError("instance_norm_relu_np_plus needs a CUDA fp32 tensor (no CPU fallback)")
        n, c, h, w = x.shape
        xc = x.contiguous()
        wt = None if weight is None else weight.detach().to(torch.float32).contiguous()
        bs = None if bias is None else bias.detach().to(torch.float32).contiguous()
        al = alpha.reshape(n, c).to(torch.float32).contiguous()
        ed = eps_draw.reshape(n, c).to(torch.float32).contiguous()
        y = torch.empty_like(xc)
        out = torch.empty_like(xc)
        mean = torch.empty((n, c), device=x.device, dtype=torch.float32)
        invstd = torch.empty((n, c), device=x.device, dtype=torch.float32)
        np_mean = torch.empty((n, c), device=x.device, dtype=torch.float32)
        psum = torch.empty((n, c), device=x.device, dtype=torch.float64)
        ws_bytes = lib.mrfp_npplus_presummed_ws_bytes(n, c)
        ws = _lib.scratch(x.device, ws_bytes, "npplus_pre")
        st = _stream_ptr(xc)
        with torch.cuda.device(x.device):
            rc = lib.mrfp_instnorm_fwd_f32(xc.data_ptr(), None if wt is None else wt.data_ptr(), None if bs is None else bs.data_ptr(),
                                           y.data_ptr(), mean.data_ptr(), invstd.data_ptr(), psum.data_ptr(), n, c, h * w,
                                           float(eps), 1, st)
            _lib.check(rc, "mrfp_instnorm_fwd_f32")
            rc = lib.mrfp_npplus_fwd_presummed_f32(y.data_ptr(), psum.data_ptr(), al.data_ptr(), ed.data_ptr(), out.data_ptr(),
                                                   np_mean.data_ptr(), None, ws.data_ptr(), ws_bytes, n, c, h * w, st)
            _lib.check(rc, "mrfp_npplus_fwd_presummed_f32")
        ctx.save_for_backward(xc, wt, bs, mean, invstd, al, ed, np_mean)
        ctx.has_affine = (weight is not None, bias is not None)
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        x, wt, bs, mean, invstd, al, ed, np_mean = ctx.saved_tensors
        n, c, h, w = x.shape
        g = g_out.contiguous()
        gx = torch.empty_like(x)
        dg = torch.empty((n, c), device=x.device, dtype=torch.float32)
        db = torch.empty((n, c), device=x.device, dtype=torch.float32)
        ws = _lib.scratch(x.device, n * c * 16, "in_np_bwd")
        with torch.cuda.device(x.device):
            rc = lib.mrfp_instnorm_bwd_np_f32(g.data_ptr(), x.data_ptr(), None if wt is None else wt.data_ptr(),
                                              None if bs is None else bs.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                              al.data_ptr(), ed.data_ptr(), np_mean.data_ptr(), ws.data_ptr(), n * c * 16,
                                              gx.data_ptr(), dg.data_ptr(), db.data_ptr(), n, c, h * w, 1, _stream_ptr(x))
        _lib.check(rc, "mrfp_instnorm_bwd_np_f32")
        gw = dg.sum(0) if ctx.has_affine[0] and ctx.needs_input_grad[1] else None
        gb = db.sum(0) if ctx.has_affine[1] and ctx.needs_input_grad[2] else None
        return gx, gw, gb, None, None, None


def instance_norm_relu_np_plus(x, weight, bias, eps, alpha, eps_draw):
    """NP+(relu(F.instance_norm(x, weight=weight, bias=bias, eps=eps))) with the two NP+ draws injected
    (alpha ~ N(1, .75), eps_draw ~ N(0, .75), shape (N, C[, 1, 1]))."""
    return _InstNormReluNpFn.apply(x, weight, bias, eps, alpha, eps_draw)


def module_instance_norm_relu_np_plus(module: torch.nn.InstanceNorm2d, x, alpha, eps_draw):
    if module.track_running_stats:
        raise _lib.MrfpError("instance_norm_relu: track_running_stats=True is not supported")
    return instance_norm_relu_np_plus(x, module.weight, module.bias, module.eps, alpha, eps_draw)
