"""CPU port of the reference's MRFP path in the reference's own terms (torch ops on the host cores).

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/mrfp_oracle.py for the rules): used by bench.py's
`cpu_baseline` leg and `--impl reference` arm — /root/reference itself is a Python package that cannot travel
to the GPU box — and by tests as a second opinion with autograd gradients at sizes numpy is too slow for.

Each function restates the cited span with the same torch operators the reference calls, so its CPU cost is
the reference's CPU cost:
  np_plus       /root/reference/deepv3.py:268-277  (draws injected instead of torch.normal)
  hrfp_chain    /root/reference/deepv3.py:320-327
  make_layers   /root/reference/deepv3.py:221-237, init network/mynn.py:57-74
Checked against the reference-generated fixtures in tests/test_oracle.py::test_torch_port_*.
"""
import math

import torch
import torch.nn.functional as F

LAYERS = ((64, 64, 1), (64, 64, 1), (64, 128, 2), (128, 256, 2), (256, 128, 1), (128, 64, 1), (64, 64, 2), (64, 64, 2))


def np_plus(feat, alpha, eps):
    feat_mean = feat.mean((2, 3), keepdim=True)                         # :269
    mean_diff = torch.std(feat_mean, 0, keepdim=True)                   # :272
    mean_scale = mean_diff / mean_diff.max() * 1.5                      # :273
    beta = 1 + eps * mean_scale                                         # :275
    return alpha * feat - alpha * feat_mean + beta * feat_mean          # :276


def make_layers(weights=None, gammas=None):
    convs, bns = [], []
    for k, (cin, cout, dil) in enumerate(LAYERS):
        c = torch.nn.Conv2d(cin, cout, 3, stride=1, padding=dil, dilation=dil).requires_grad_(False)
        b = torch.nn.BatchNorm2d(cout).requires_grad_(False)
        torch.nn.init.kaiming_normal_(c.weight, nonlinearity="relu"); c.bias.data.zero_()
        torch.nn.init.normal_(b.weight, mean=0.0, std=0.5); b.bias.data.zero_()
        if weights is not None:
            with torch.no_grad():
                c.weight.copy_(torch.as_tensor(weights[k])); b.weight.copy_(torch.as_tensor(gammas[k]))
        convs.append(c); bns.append(b)
    return convs, bns


class _NearestExactAdjoint(torch.autograd.Function):
    """F.interpolate(mode='nearest') with ATen's forward index rule and the EXACT adjoint as backward.

    Needed on CUDA only: ATen's CUDA `upsample_nearest2d_backward` is not the adjoint of its own forward for
    scale_factor=1.2 (the two x1.2 stages of deepv3.py:321-322).  The forward picks src = floorf(dst * float(1/1.2)); the
    backward re-derives the replica range of a source pixel as [ceilf(src * 1.2f), ceilf((src+1) * 1.2f)) and the two
    float roundings disagree at every fifth pixel: <L x, v> and <x, L^T v> differ by 20-40 % (tools/adjoint_nearest.py on
    the B200 box, torch 2.11; scale factors 1.205 / 0.838 / 0.798 and the size= calls are consistent, and so is the CPU
    kernel for every case).  The reference's CPU autograd, the fixtures under tests/golden/ and this repo's kernels all
    implement the adjoint of the forward."""

    @staticmethod
    def forward(ctx, x, idx_h, idx_w):
        ctx.save_for_backward(idx_h, idx_w)
        ctx.in_hw = (x.shape[2], x.shape[3])
        return x.index_select(2, idx_h).index_select(3, idx_w)

    @staticmethod
    def backward(ctx, g):
        idx_h, idx_w = ctx.saved_tensors
        ih, iw = ctx.in_hw
        t = g.new_zeros(g.shape[0], g.shape[1], ih, g.shape[3]).index_add_(2, idx_h, g)
        return g.new_zeros(g.shape[0], g.shape[1], ih, iw).index_add_(3, idx_w, t), None, None


def _nearest_index(in_size, out_size, scale_factor, device):
    """ATen's forward rule (upsample_nearest2d): src = min(floorf(dst * scale), in - 1), scale a float32."""
    scale = torch.tensor(1.0 / scale_factor if scale_factor else in_size / out_size, dtype=torch.float64).to(torch.float32)
    if not scale_factor:
        scale = torch.tensor(in_size, dtype=torch.float32) / torch.tensor(out_size, dtype=torch.float32)
    dst = torch.arange(out_size, dtype=torch.float32)
    return torch.clamp(torch.floor(dst * scale).to(torch.int64), max=in_size - 1).to(device)


def nearest(x, scale_factor=None, size=None, exact_adjoint=False):
    """The reference's F.interpolate call (default mode='nearest'); exact_adjoint: see _NearestExactAdjoint."""
    if not exact_adjoint:
        return F.interpolate(x, scale_factor=scale_factor, size=size)
    ih, iw = x.shape[2], x.shape[3]
    if scale_factor is not None:
        oh, ow = int(math.floor(ih * scale_factor[0])), int(math.floor(iw * scale_factor[1]))
        idx_h, idx_w = _nearest_index(ih, oh, scale_factor[0], x.device), _nearest_index(iw, ow, scale_factor[1], x.device)
    else:
        oh, ow = size
        idx_h, idx_w = _nearest_index(ih, oh, None, x.device), _nearest_index(iw, ow, None, x.device)
    return _NearestExactAdjoint.apply(x, idx_h, idx_w)


def hrfp_chain(convs, bns, xp, h, w, exact_adjoint=False):
    """deepv3.py:320-327.  exact_adjoint=True replaces only the BACKWARD of the nearest resamples (CUDA runs)."""
    ea = exact_adjoint
    o = F.relu(bns[0](nearest(convs[0](xp), scale_factor=(1.205, 1.205), exact_adjoint=ea)))
    o = F.relu(bns[1](nearest(convs[1](o), scale_factor=(1.2, 1.2), exact_adjoint=ea)))
    o = F.relu(bns[2](nearest(convs[2](o), scale_factor=(1.2, 1.2), exact_adjoint=ea)))
    dec = F.relu(bns[3](nearest(convs[3](o), size=(int(h / 2), int(w / 2)), exact_adjoint=ea)))
    o = F.relu(bns[4](nearest(convs[4](dec), size=(int(h / 2), int(w / 2)), exact_adjoint=ea)))
    o = F.relu(bns[5](nearest(convs[5](o), scale_factor=(0.838, 0.838), exact_adjoint=ea)))
    o = F.relu(bns[6](nearest(convs[6](o), scale_factor=(0.798, 0.798), exact_adjoint=ea)))
    o = F.relu(bns[7](nearest(convs[7](o), size=(math.ceil(h / 4), math.ceil(w / 4)), exact_adjoint=ea)))
    return o, dec


def mrfp_step(convs, bns, xp, feat2, dec1, draws, grads, h, w, final2=None):
    """One forward+backward pass of the whole MRFP path: both NP+ calls, the HRFP chain, the HRFP and HRFP+ adds and —
    with `final2` (deepv3.py:219-220, :359) — the 1x1 classifier that consumes the HRFP+ sum.
    `dec1` is the decoder feature before the reference's Upsample (deepv3.py:356, mynn.py:114-119); a tensor that
    already has the (h/2, w/2) size passes through the interpolation unchanged.  grads = (g_x, g_tail, g_feat2) with
    g_tail the gradient of the last tensor of the tail (dec2 with final2, the HRFP+ sum without)."""
    (a1, e1), (a2, e2) = draws
    g_x, g_d1, g_f2 = grads
    xp = xp.detach().requires_grad_(True)
    feat2 = feat2.detach().requires_grad_(True)
    dec1 = dec1.detach().requires_grad_(True)
    x = np_plus(xp, a1, e1)                                              # :318
    ocout, dec = hrfp_chain(convs, bns, xp, h, w)                        # :320-327
    x = ocout + x                                                        # :330
    y2 = np_plus(feat2, a2, e2)                                          # :335
    dec1_up = F.interpolate(dec1, size=(int(h / 2), int(w / 2)), mode="bilinear", align_corners=True)   # :356
    d1 = torch.add(dec, dec1_up)                                         # :357
    tail = final2(d1) if final2 is not None else d1                      # :359
    torch.autograd.backward([x, tail, y2], [g_x, g_d1, g_f2])
    return x.detach(), tail.detach(), y2.detach(), xp.grad, feat2.grad, dec1.grad
