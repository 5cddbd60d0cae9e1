"""CPU oracle for the MRFP hot path (NP+ and HRFP/HRFP+), numpy float64/float32.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`mrfp_b200/`) may import this
module; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs use it, and only as the checker / the timed CPU baseline.

It is an explicit restatement (no torch, no autograd) of the reference algorithm:

* NP+            /root/reference/deepv3.py:268-277   (`Normalization_Perturbation_Plus`)
* HRFP chain     /root/reference/deepv3.py:320-327   (conv -> nearest resample -> BN(train) -> ReLU, x8)
* HRFP add       /root/reference/deepv3.py:329-330
* HRFP+ add      /root/reference/deepv3.py:355-357
* HRFP init      /root/reference/network/mynn.py:57-74
* layer shapes   /root/reference/deepv3.py:221-237

The arithmetic of the reference lives in PyTorch (pinned 1.12.1 in SDG.yml:153; 2.11.0 here):
`mean`, `std` (unbiased), `max`, `Conv2d`, `F.interpolate(mode='nearest')`, `BatchNorm2d`
(train mode), `relu`.  Their published semantics are restated here in numpy.

Pinning: the reference ships no tests / golden vectors for this path, so
the oracle is pinned against outputs of the reference itself, generated in the build
container by `tests/golden/make_golden.py` (imports /root/reference through a stub shim) and
committed as `tests/golden/*.npz`; `tests/test_oracle.py` checks every function below
against them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# HRFP layer table — deepv3.py:221-237  (cin, cout, dilation); kernel 3, stride 1, pad = dilation
# --------------------------------------------------------------------------------------
HRFP_LAYERS: Tuple[Tuple[int, int, int], ...] = (
    (64, 64, 1),     # OClayer1     deepv3.py:221
    (64, 64, 1),     # OClayer2     deepv3.py:223
    (64, 128, 2),    # OClayer3     deepv3.py:225
    (128, 256, 2),   # OClayer4     deepv3.py:227
    (256, 128, 1),   # OCdeclayer1  deepv3.py:230
    (128, 64, 1),    # OCdeclayer2  deepv3.py:232
    (64, 64, 2),     # OCdeclayer3  deepv3.py:234
    (64, 64, 2),     # OCdeclayer4  deepv3.py:236
)
HRFP_CONV_NAMES = ("OClayer1", "OClayer2", "OClayer3", "OClayer4",
                   "OCdeclayer1", "OCdeclayer2", "OCdeclayer3", "OCdeclayer4")
HRFP_BN_NAMES = ("OC1_bn", "OC2_bn", "OC3_bn", "OC4_bn",
                 "OC1_decbn", "OC2_decbn", "OC3_decbn", "OC4_decbn")
BN_EPS = 1e-5          # nn.BatchNorm2d default, deepv3.py:222
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# nearest-neighbour resampling geometry — F.interpolate(mode='nearest'), deepv3.py:320-327
# --------------------------------------------------------------------------------------
def nearest_out_size(in_size: int, scale_factor: float) -> int:
    """Output size in scale_factor mode: floor(in * scale_factor) evaluated in double."""
    return int(math.floor(float(in_size) * float(scale_factor)))


def nearest_src_index(in_size: int, out_size: int, scale_factor: Optional[float] = None) -> np.ndarray:
    """dst -> src index of ATen's `upsample_nearest2d`.

    scale = float32(1/scale_factor) when a scale factor was given (recompute_scale_factor=None),
    else float32(in)/float32(out); src = min(int(floorf(dst * scale)), in - 1), all in float32.
    The CPU kernel special-cases out==in (identity) and out==2*in (dst>>1); both agree with the rule
    whenever no scale factor is supplied.
    """
    if scale_factor is not None and scale_factor > 0:
        scale = np.float32(1.0 / float(scale_factor))
    else:
        scale = np.float32(in_size) / np.float32(out_size)
    dst = np.arange(out_size, dtype=np.float32)
    src = np.floor(dst * scale).astype(np.int64)
    return np.minimum(src, in_size - 1)


@dataclass
class HrfpStage:
    cin: int
    cout: int
    dil: int
    conv_h: int          # conv runs at this resolution (input and output of the conv)
    conv_w: int
    out_h: int           # resolution after the nearest resample
    out_w: int
    idx_h: np.ndarray    # (out_h,) dst row -> src row
    idx_w: np.ndarray    # (out_w,)


def hrfp_geometry(h: int, w: int, xh: int, xw: int,
                  layers: Sequence[Tuple[int, int, int]] = HRFP_LAYERS) -> List[HrfpStage]:
    """Sizes of the 8 stages for an (h, w) image whose stem feature `xp` is (xh, xw).

    deepv3.py:320-327: scale factors 1.205, 1.2, 1.2, size (h/2,w/2), size (h/2,w/2),
    0.838, 0.798, size (ceil(h/4), ceil(w/4)).
    """
    spec = [("sf", 1.205), ("sf", 1.2), ("sf", 1.2), ("size", (int(h / 2), int(w / 2))),
            ("size", (int(h / 2), int(w / 2))), ("sf", 0.838), ("sf", 0.798),
            ("size", (math.ceil(h / 4), math.ceil(w / 4)))]
    stages = []
    ch, cw = xh, xw
    for (cin, cout, dil), (mode, arg) in zip(layers, spec):
        if mode == "sf":
            oh, ow = nearest_out_size(ch, arg), nearest_out_size(cw, arg)
            ih, iw = nearest_src_index(ch, oh, arg), nearest_src_index(cw, ow, arg)
        else:
            oh, ow = arg
            ih, iw = nearest_src_index(ch, oh), nearest_src_index(cw, ow)
        stages.append(HrfpStage(cin, cout, dil, ch, cw, oh, ow, ih, iw))
        ch, cw = oh, ow
    return stages


# --------------------------------------------------------------------------------------
# NP+ — deepv3.py:268-277
# --------------------------------------------------------------------------------------
def np_plus_stats(mean: np.ndarray, eps_draw: np.ndarray):
    """(N,C) plane means -> d (C,), dmax, s (C,), beta (N,C).  deepv3.py:272-275."""
    n = mean.shape[0]
    with np.errstate(all="ignore"):
        mbar = mean.mean(0, keepdims=True)
        d = np.sqrt(((mean - mbar) ** 2).sum(0) / (n - 1))        # torch.std(.,0): unbiased
        dmax = d.max() if not np.isnan(d).any() else np.float64(np.nan)   # Tensor.max propagates NaN
        s = d / dmax * 1.5
        beta = 1.0 + eps_draw * s[None, :]
    return d, dmax, s, beta


def np_plus_forward(feat: np.ndarray, alpha_draw: np.ndarray, eps_draw: np.ndarray):
    """feat (N,C,H,W); alpha_draw = torch.normal(1,.75) and eps_draw = torch.normal(0,.75), both (N,C).

    Returns (out, mean (N,C), beta (N,C)).  deepv3.py:269-277.
    """
    feat = np.asarray(feat)
    mean = feat.mean((2, 3))                                      # :269
    _, _, _, beta = np_plus_stats(mean, eps_draw)                 # :272-275
    a = alpha_draw[:, :, None, None]
    m = mean[:, :, None, None]
    with np.errstate(all="ignore"):
        out = a * feat - a * m + beta[:, :, None, None] * m       # :276
    return out, mean, beta


def np_plus_backward(gout: np.ndarray, alpha_draw: np.ndarray, eps_draw: np.ndarray, mean: np.ndarray):
    """Closed-form input gradient of NP+ (autograd flows through std and max; nothing is detached).

    G = sum_hw g;  dL/ds[c] = sum_n eps*m*G;  dL/dd[c] = 1.5/dmax*dL/ds[c] - [c==c*] * sum_c' dL/ds[c']*1.5*d[c']/dmax^2;
    dL/dm = (beta-alpha)*G + dL/dd[c]*(m-mbar)/((N-1)*d[c])  (second term := 0 where d[c]==0, as torch's
    std_backward does);   gin = alpha*g + dL/dm/(H*W).
    """
    n, c, hh, ww = gout.shape
    d, dmax, s, beta = np_plus_stats(mean, eps_draw)
    G = gout.sum((2, 3))
    with np.errstate(all="ignore"):
        dL_ds = (eps_draw * mean * G).sum(0)
        dL_dd = 1.5 / dmax * dL_ds
        cstar = int(np.argmax(d)) if not np.isnan(d).any() else 0
        dL_dd[cstar] -= (dL_ds * 1.5 * d).sum() / (dmax * dmax)
        mbar = mean.mean(0, keepdims=True)
        # torch's std_backward zero-fills the gradient where std == 0 (a channel whose plane means coincide)
        dd_dm = np.where(d[None, :] == 0, 0.0, (mean - mbar) / ((n - 1) * d[None, :]))
        dL_dm = (beta - alpha_draw) * G + dL_dd[None, :] * dd_dm
        gin = alpha_draw[:, :, None, None] * gout + (dL_dm / (hh * ww))[:, :, None, None]
    return gin


# --------------------------------------------------------------------------------------
# HRFP building blocks
# --------------------------------------------------------------------------------------
def conv3x3(x: np.ndarray, wgt: np.ndarray, dil: int, bias: Optional[np.ndarray] = None) -> np.ndarray:
    """nn.Conv2d(k=3, stride=1, padding=dil, dilation=dil) — cross-correlation, zero padding."""
    n, cin, hh, ww = x.shape
    cout = wgt.shape[0]
    xp = np.zeros((n, cin, hh + 2 * dil, ww + 2 * dil), dtype=x.dtype)
    xp[:, :, dil:dil + hh, dil:dil + ww] = x
    y = np.zeros((n, cout, hh, ww), dtype=x.dtype)
    for ky in range(3):
        for kx in range(3):
            y += np.einsum("oc,nchw->nohw", wgt[:, :, ky, kx],
                           xp[:, :, ky * dil:ky * dil + hh, kx * dil:kx * dil + ww], optimize=True)
    if bias is not None:
        y += bias[None, :, None, None]
    return y


def conv3x3_dgrad(gy: np.ndarray, wgt: np.ndarray, dil: int) -> np.ndarray:
    """Input gradient of conv3x3: correlation of gy with the 180-degree-rotated, io-transposed kernel."""
    n, cout, hh, ww = gy.shape
    cin = wgt.shape[1]
    gp = np.zeros((n, cout, hh + 2 * dil, ww + 2 * dil), dtype=gy.dtype)
    gp[:, :, dil:dil + hh, dil:dil + ww] = gy
    gx = np.zeros((n, cin, hh, ww), dtype=gy.dtype)
    for ky in range(3):
        for kx in range(3):
            gx += np.einsum("oc,nohw->nchw", wgt[:, :, 2 - ky, 2 - kx],
                            gp[:, :, ky * dil:ky * dil + hh, kx * dil:kx * dil + ww], optimize=True)
    return gx


def resample(y: np.ndarray, idx_h: np.ndarray, idx_w: np.ndarray) -> np.ndarray:
    return y[:, :, idx_h][:, :, :, idx_w]


def resample_bwd(gr: np.ndarray, idx_h: np.ndarray, idx_w: np.ndarray, in_h: int, in_w: int) -> np.ndarray:
    """Adjoint of `resample`: sum of the gradients of every replica (zero where never sampled)."""
    n, c = gr.shape[:2]
    tmp = np.zeros((n, c, in_h, gr.shape[3]), dtype=gr.dtype)
    np.add.at(tmp, (slice(None), slice(None), idx_h), gr)
    out = np.zeros((n, c, in_h, in_w), dtype=gr.dtype)
    np.add.at(out, (slice(None), slice(None), slice(None), idx_w), tmp)
    return out


def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16 precision (returned as float64)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32).astype(np.float64)


def hrfp_forward(xp: np.ndarray, weights: Sequence[np.ndarray], gammas: Sequence[np.ndarray],
                 h: int, w: int, betas: Optional[Sequence[np.ndarray]] = None,
                 biases: Optional[Sequence[np.ndarray]] = None,
                 layers: Sequence[Tuple[int, int, int]] = HRFP_LAYERS, quant=None):
    """deepv3.py:320-327.  Returns (OCout, OCout_dec, saved) with saved = per-stage dict for backward
    and for the BN running-stat side effect (a-8): batch mean, biased var, element count.

    `quant` (e.g. `round_bf16`) models a reduced-precision STORAGE format: it is applied where the
    tensor-core CUDA path stores bf16 (stem feature, weights, each conv output, each activation); the
    arithmetic stays exact and the BN statistics are those of the stored (rounded) conv output, as on the device."""
    stages = hrfp_geometry(h, w, xp.shape[2], xp.shape[3], layers)
    q = quant if quant is not None else (lambda t: t)
    a = q(xp)
    saved = []
    ocout_dec = None
    for k, st in enumerate(stages):
        y_acc = conv3x3(a, q(weights[k]), st.dil, None if biases is None else biases[k])
        r = resample(q(y_acc), st.idx_h, st.idx_w)               # statistics of the STORED conv output, as on the device
        mu = r.mean((0, 2, 3))
        var = r.var((0, 2, 3))                                   # biased, used for normalisation
        invstd = 1.0 / np.sqrt(var + BN_EPS)
        xhat = (r - mu[None, :, None, None]) * invstd[None, :, None, None]
        z = xhat * gammas[k][None, :, None, None]
        if betas is not None:
            z = z + betas[k][None, :, None, None]
        a_next = np.maximum(z, 0)
        saved.append(dict(stage=st, a_in=a, xhat=xhat, z=z, invstd=invstd, mean=mu, var=var,
                          count=r.shape[0] * r.shape[2] * r.shape[3]))
        if k == 3:
            ocout_dec = a_next               # the fp32 NCHW outputs are written unrounded
        a = q(a_next) if k < len(stages) - 1 else a_next
    return a, ocout_dec, saved


def hrfp_backward(g_ocout: Optional[np.ndarray], g_ocout_dec: Optional[np.ndarray],
                  weights: Sequence[np.ndarray], gammas: Sequence[np.ndarray], saved, quant=None) -> np.ndarray:
    """Input gradient (wrt xp) of the chain; no weight / gamma / beta gradients (frozen, deepv3.py:221-237).
    `quant`: storage rounding of the incoming gradients, of dY and of each dgrad output (see hrfp_forward)."""
    q = quant if quant is not None else (lambda t: t)
    ga = None if g_ocout is None else q(g_ocout)
    for k in range(7, -1, -1):
        sv = saved[k]
        st: HrfpStage = sv["stage"]
        if k == 3:
            if ga is None:
                ga = None if g_ocout_dec is None else q(g_ocout_dec)
            elif g_ocout_dec is not None:
                ga = q(ga + g_ocout_dec)
        if ga is None:
            continue
        dz = ga * (sv["z"] > 0)
        dxh = dz * gammas[k][None, :, None, None]
        m1 = dxh.mean((0, 2, 3))[None, :, None, None]
        m2 = (dxh * sv["xhat"]).mean((0, 2, 3))[None, :, None, None]
        dr = sv["invstd"][None, :, None, None] * (dxh - m1 - sv["xhat"] * m2)
        dy = q(resample_bwd(dr, st.idx_h, st.idx_w, st.conv_h, st.conv_w))
        ga = q(conv3x3_dgrad(dy, q(weights[k]), st.dil))
    return ga


def bn_running_update(running_mean, running_var, batch_mean, batch_var_biased, count, momentum=BN_MOMENTUM):
    """nn.BatchNorm2d train-mode buffer update: running_var uses the unbiased batch variance."""
    unbiased = batch_var_biased * (count / (count - 1.0))
    return ((1 - momentum) * running_mean + momentum * batch_mean,
            (1 - momentum) * running_var + momentum * unbiased)


def hrfp_init_std(cin: int) -> float:
    """kaiming_normal_(nonlinearity='relu') on a (cout,cin,3,3) weight: std = sqrt(2 / (9*cin)). mynn.py:64."""
    return math.sqrt(2.0 / (9.0 * cin))


# --------------------------------------------------------------------------------------
# HRFP+ skip — deepv3.py:355-357 (bilinear align_corners upsample, network/mynn.py:114-119)
# --------------------------------------------------------------------------------------
def bilinear_align_corners(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    n, c, ih, iw = x.shape

    def coords(i_n, o_n):
        if o_n == 1:
            src = np.zeros(1)
        else:
            src = np.arange(o_n, dtype=np.float64) * ((i_n - 1) / (o_n - 1))
        i0 = np.minimum(np.floor(src).astype(np.int64), i_n - 1)
        i1 = np.minimum(i0 + 1, i_n - 1)
        return i0, i1, src - i0

    y0, y1, fy = coords(ih, out_h)
    x0, x1, fx = coords(iw, out_w)
    top = x[:, :, y0][:, :, :, x0] * (1 - fx) + x[:, :, y0][:, :, :, x1] * fx
    bot = x[:, :, y1][:, :, :, x0] * (1 - fx) + x[:, :, y1][:, :, :, x1] * fx
    return top * (1 - fy)[:, None] + bot * fy[:, None]


def hrfp_plus_add(dec1: np.ndarray, ocout_dec: np.ndarray) -> np.ndarray:
    """deepv3.py:356-357."""
    return bilinear_align_corners(dec1, ocout_dec.shape[2], ocout_dec.shape[3]) + ocout_dec


# --------------------------------------------------------------------------------------
# InstanceNorm2d(affine=True) + ReLU of the trunk — network/Resnet.py:176-178 + :218-225 (last
# Bottleneck of layer1 / layer2, iw == 4) and :534-536 + :596-598 (stem, wt_layer[2] == 4).  nn.InstanceNorm2d
# semantics: per-(n, c) plane, biased variance, eps inside the square root, no running statistics.
# --------------------------------------------------------------------------------------
IN_EPS = 1e-5          # nn.InstanceNorm2d default


def instance_norm_relu_forward(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, eps: float = IN_EPS, relu: bool = True):
    """Returns (y, mean (N,C), invstd (N,C), psum (N,C) = sum_hw y)."""
    x = np.asarray(x, dtype=np.float64)
    mean = x.mean(axis=(2, 3))
    var = ((x - mean[:, :, None, None]) ** 2).mean(axis=(2, 3))
    invstd = 1.0 / np.sqrt(var + eps)
    y = (x - mean[:, :, None, None]) * invstd[:, :, None, None] * np.asarray(gamma, np.float64)[None, :, None, None] \
        + np.asarray(beta, np.float64)[None, :, None, None]
    if relu:
        y = np.maximum(y, 0.0)
    return y, mean, invstd, y.sum(axis=(2, 3))


def instance_norm_relu_backward(gy: np.ndarray, x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, eps: float = IN_EPS,
                                relu: bool = True):
    """Closed-form backward of relu(IN(x)): returns (gx, d_gamma (C), d_beta (C)).  ReLU passes gradient where y > 0
    (threshold_backward)."""
    x = np.asarray(x, dtype=np.float64)
    g = np.asarray(gy, dtype=np.float64).copy()
    gm = np.asarray(gamma, np.float64)[None, :, None, None]
    mean = x.mean(axis=(2, 3), keepdims=True)
    var = ((x - mean) ** 2).mean(axis=(2, 3), keepdims=True)
    invstd = 1.0 / np.sqrt(var + eps)
    xh = (x - mean) * invstd
    if relu:
        g[(xh * gm + np.asarray(beta, np.float64)[None, :, None, None]) <= 0.0] = 0.0
    m1 = g.mean(axis=(2, 3), keepdims=True)
    m2 = (g * xh).mean(axis=(2, 3), keepdims=True)
    gx = gm * invstd * (g - m1 - xh * m2)
    return gx, (g * xh).sum(axis=(0, 2, 3)), g.sum(axis=(0, 2, 3))
