"""Error report of the bf16 tensor-core HRFP chain: vs the fp32 reference fixture and vs the bf16-storage oracle."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import mrfp_oracle as O
from tests.common import GOLDEN, make_hrfp_params, make_feat
from tests.test_hrfp_gpu import _run

def rel(got, ref):
    ref = np.asarray(ref, np.float64); d = got.astype(np.float64) - ref
    return np.abs(d).max() / np.abs(ref).max(), np.sqrt((d * d).sum() / (ref * ref).sum())

g = np.load(os.path.join(GOLDEN, "hrfp.npz"))
for tag in ("sq", "rect"):
    n, h, w, seed = [int(v) for v in g[f"{tag}_meta"]]
    xh, xw = math.ceil(h / 4), math.ceil(w / 4)
    ws, gs = make_hrfp_params(seed)
    xp = make_feat(seed + 50, (n, 64, xh, xw))
    rng = np.random.default_rng(seed + 70)
    g1 = rng.standard_normal((n, 64, xh, xw)).astype(np.float32)
    g2 = rng.standard_normal((n, 256, h // 2, w // 2)).astype(np.float32)
    ws64 = [a.astype(np.float64) for a in ws]; gs64 = [a.astype(np.float64) for a in gs]
    qo, qd, qs = O.hrfp_forward(xp.astype(np.float64), ws64, gs64, h, w, quant=O.round_bf16)
    qg = O.hrfp_backward(g1.astype(np.float64), g2.astype(np.float64), ws64, gs64, qs, quant=O.round_bf16)
    for mode in (0, 2):
        out, dec, gx, _, _ = _run(xp, ws, gs, h, w, mode, g1, g2)
        print(tag, "mode", mode, "vs fp32 reference: out max/L2 %.2e %.2e | dec %.2e %.2e | gx %.2e %.2e" % (
            *rel(out, g[f"{tag}_ocout"]), *rel(dec, g[f"{tag}_ocout_dec"].astype(np.float32)), *rel(gx, g[f"{tag}_gx_both"])))
        if mode == 2:
            print(tag, "mode", mode, "vs bf16-storage oracle: out max/L2 %.2e %.2e | dec %.2e %.2e | gx %.2e %.2e" % (
                *rel(out, qo), *rel(dec, qd), *rel(gx, qg)))
    print(tag, "bf16-storage oracle vs fp32 reference: out %.2e %.2e | gx %.2e %.2e" % (*rel(qo, g[f"{tag}_ocout"]), *rel(qg, g[f"{tag}_gx_both"])))
