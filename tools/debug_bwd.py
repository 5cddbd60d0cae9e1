"""Localises a backward discrepancy of the HRFP chain: size scan x gradient path, fp32 CUDA-core mode vs torch fp64."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import torch_port as TP
from tests.common import make_feat, make_hrfp_params
from mrfp_b200.hrfp import hrfp_chain

torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
def l2(a, b):
    b = b.detach().double(); return float(((a.detach().double() - b).norm() / b.norm()).item())
mode = int(os.environ.get("MODE", "0"))
for h in (96, 192, 384, 512, 768):
    n, xh = 2, h // 4
    ws, gs = make_hrfp_params(11)
    xp = torch.from_numpy(make_feat(12, (n, 64, xh, xh))).cuda()
    g1 = torch.randn(n, 64, xh, xh, device="cuda"); g2 = torch.randn(n, 256, h // 2, h // 2, device="cuda")
    c64, b64 = TP.make_layers(ws, gs); c64 = [c.to("cuda", torch.float64) for c in c64]; b64 = [b.to("cuda", torch.float64) for b in b64]
    for which in ("out", "dec", "both"):
        x64 = xp.double().requires_grad_(True)
        o, d = TP.hrfp_chain(c64, b64, x64, h, h, exact_adjoint=os.environ.get('EXACT','1')=='1')
        outs, grads = [], []
        if which in ("out", "both"): outs.append(o); grads.append(g1.double())
        if which in ("dec", "both"): outs.append(d); grads.append(g2.double())
        torch.autograd.backward(outs, grads)
        cc, bb = TP.make_layers(ws, gs); cc = [c.cuda() for c in cc]; bb = [b.cuda() for b in bb]
        xa = xp.clone().requires_grad_(True)
        oo, dd = hrfp_chain(xa, cc, bb, h, h, math_mode=mode, want_out=which != "dec", want_dec=which != "out")
        outs, grads = [], []
        if which in ("out", "both"): outs.append(oo); grads.append(g1)
        if which in ("dec", "both"): outs.append(dd); grads.append(g2)
        torch.autograd.backward(outs, grads)
        e = (xa.grad.double() - x64.grad); proj = float((e.flatten() @ x64.grad.flatten()) / (x64.grad.flatten() @ x64.grad.flatten()))
        print(f"h={h} mode={mode} path={which}: gx l2 {l2(xa.grad, x64.grad):.3e} proj {proj:+.3e} |ref| {float(x64.grad.norm()):.3e} |ours| {float(xa.grad.norm()):.3e}", flush=True)
