import os, sys, random, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from tests.common import GOLDEN, make_hrfp_params, make_draws, fill_state_dict
from tests.test_model import _criterion
from mrfp_b200 import model as M, npplus, hrfp
from tools.bench_hrfp import eager_chain
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False

def build():
    m = M.MRFPPlus(19, criterion=_criterion(), math_mode=0)
    fill_state_dict(m, 77)
    return m.cuda().train()

def run(m, eager):
    n, hh, ww = 2, 64, 64
    rng = np.random.default_rng(78)
    x = torch.from_numpy(rng.uniform(0, 255, (n, 3, hh, ww)).astype(np.float32)).cuda()
    gts = torch.from_numpy(rng.integers(0, 19, (n, hh, ww)).astype(np.int64))
    gts[torch.from_numpy(rng.uniform(size=(n, hh, ww)) < 0.05)] = 255
    gts = gts.cuda()
    ws, gs = make_hrfp_params(79)
    draws = [make_draws(80, n, 64), make_draws(81, n, 256)]
    def fake_reinit():
        convs, bns = m.hrfp_modules()
        with torch.no_grad():
            for k in range(8):
                convs[k].weight.copy_(torch.from_numpy(ws[k])); convs[k].bias.zero_()
                bns[k].weight.copy_(torch.from_numpy(gs[k])); bns[k].bias.zero_()
    m.reinit_hrfp = fake_reinit
    def fake_draws(feat):
        a, e = draws.pop(0)
        return torch.from_numpy(a).to(feat.device).view(n, -1, 1, 1), torch.from_numpy(e).to(feat.device).view(n, -1, 1, 1)
    if eager:
        def np_eager(feat):
            a, e = fake_draws(feat)
            mean = feat.mean((2, 3), keepdim=True)
            d = torch.std(mean, 0, keepdim=True)
            s = d / d.max() * 1.5
            b = 1 + e * s
            return a * feat - a * mean + b * mean
        m.Normalization_Perturbation_Plus = np_eager
        def stem(xp, h, w, training, p, p2, p3):
            x = xp
            if p2 < .5: x = np_eager(xp)
            convs, bns = m.hrfp_modules()
            o, d = eager_chain(convs, bns, xp, h, w)
            if p < .5: x = o + x
            return x, d
        m.mrfp_stem = stem
        hrfp_add = hrfp.hrfp_plus_add
        hrfp.hrfp_plus_add = lambda a, b: a + b
    orig = npplus.draw_np_plus_factors
    npplus.draw_np_plus_factors = fake_draws
    random.seed(4)
    loss = m(x, gts, training=True)
    loss.backward()
    npplus.draw_np_plus_factors = orig
    return float(loss), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}

l1, g1 = run(build(), False)
l2, g2 = run(build(), True)
print("loss mine", l1, "eager", l2, "golden", float(np.load(os.path.join(GOLDEN, "full_model.npz"))["loss"]))
worst = []
for k in g1:
    d = (g1[k] - g2[k]).abs().max().item() / max(g2[k].abs().max().item(), 1e-20)
    worst.append((d, k))
worst.sort(reverse=True)
for d, k in worst[:12]: print("%.3e %s" % (d, k))
print("...")
for d, k in worst[-3:]: print("%.3e %s" % (d, k))
g = np.load(os.path.join(GOLDEN, "full_model.npz"))
for key in [k[3:] for k in g.files if k.startswith("gs_")]:
    for name, gg in (("mine", g1), ("eager", g2)):
        grad = gg[key].double().cpu()
        samp = grad.flatten()[:: max(1, grad.numel() // 64)][:64].numpy()
        print(key, name, "max diff vs golden %.3e (max ref %.3e)" % (np.abs(samp - g["gs_" + key]).max(), np.abs(g["gs_" + key]).max()))
