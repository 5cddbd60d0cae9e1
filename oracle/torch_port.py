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


def hrfp_chain(convs, bns, xp, h, w):
    o = F.relu(bns[0](F.interpolate(convs[0](xp), scale_factor=(1.205, 1.205))))
    o = F.relu(bns[1](F.interpolate(convs[1](o), scale_factor=(1.2, 1.2))))
    o = F.relu(bns[2](F.interpolate(convs[2](o), scale_factor=(1.2, 1.2))))
    dec = F.relu(bns[3](F.interpolate(convs[3](o), size=(int(h / 2), int(w / 2)))))
    o = F.relu(bns[4](F.interpolate(convs[4](dec), size=(int(h / 2), int(w / 2)))))
    o = F.relu(bns[5](F.interpolate(convs[5](o), scale_factor=(0.838, 0.838))))
    o = F.relu(bns[6](F.interpolate(convs[6](o), scale_factor=(0.798, 0.798))))
    o = F.relu(bns[7](F.interpolate(convs[7](o), size=(math.ceil(h / 4), math.ceil(w / 4)))))
    return o, dec


def mrfp_step(convs, bns, xp, feat2, dec1, draws, grads, h, w):
    """One forward+backward pass of the whole MRFP path: both NP+ calls, the HRFP chain, the HRFP and HRFP+ adds.
    `dec1` is the decoder feature before the reference's Upsample (deepv3.py:356, mynn.py:114-119); a tensor that
    already has the (h/2, w/2) size passes through the interpolation unchanged."""
    (a1, e1), (a2, e2) = draws
    g_x, g_d1, g_f2 = grads
    xp = xp.detach().requires_grad_(True)
    feat2 = feat2.detach().requires_grad_(True)
    x = np_plus(xp, a1, e1)                                              # :318
    ocout, dec = hrfp_chain(convs, bns, xp, h, w)                        # :320-327
    x = ocout + x                                                        # :330
    y2 = np_plus(feat2, a2, e2)                                          # :335
    dec1_up = F.interpolate(dec1, size=(int(h / 2), int(w / 2)), mode="bilinear", align_corners=True)   # :356
    d1 = torch.add(dec, dec1_up)                                         # :357
    torch.autograd.backward([x, d1, y2], [g_x, g_d1, g_f2])
    return x.detach(), d1.detach(), y2.detach(), xp.grad, feat2.grad
