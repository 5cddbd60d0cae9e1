"""Two passes of the bench step (for ncu captures): chain with NP+ call 1 folded in, NP+ call 2, tail through the classifier."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrfp_b200 import hrfp as H, npplus as NP
from mrfp_b200.model import init_hrfp_module
dev = "cuda"
torch.manual_seed(1)
n = 8
chans, dils = [64, 64, 64, 128, 256, 128, 64, 64, 64], [1, 1, 2, 2, 1, 1, 2, 2]
convs = [torch.nn.Conv2d(chans[k], chans[k + 1], 3, padding=dils[k], dilation=dils[k]).to(dev).requires_grad_(False) for k in range(8)]
bns = [torch.nn.BatchNorm2d(chans[k + 1]).to(dev).requires_grad_(False) for k in range(8)]
for c, b in zip(convs, bns):
    init_hrfp_module(c); init_hrfp_module(b)
final2 = torch.nn.Conv2d(256, 19, 1).to(dev)
xp = torch.relu(torch.randn(n, 64, 192, 192, device=dev))
f2 = torch.relu(torch.randn(n, 256, 192, 192, device=dev))
d1 = torch.randn(n, 256, 192, 192, device=dev)
draws = [(1 + 0.75 * torch.randn(n, c, 1, 1, device=dev), 0.75 * torch.randn(n, c, 1, 1, device=dev)) for c in (64, 256)]
g_x = torch.randn(n, 64, 192, 192, device=dev); g_d = torch.randn(n, 19, 384, 384, device=dev); g_f = torch.randn(n, 256, 192, 192, device=dev)
for _ in range(2):
    a = xp.detach().requires_grad_(True); b = f2.detach().requires_grad_(True); d = d1.detach().requires_grad_(True)
    final2.zero_grad(set_to_none=True)
    x, dec = H.hrfp_chain(a, convs, bns, 768, 768, np_draws=draws[0], math_mode=H.MATH_BF16, lazy_dec=True)
    y2 = NP.np_plus_with_draws(b, *draws[1])
    o = H.hrfp_plus_final2(d, final2, dec)
    torch.autograd.backward([x, o, y2], [g_x, g_d, g_f])
torch.cuda.synchronize()
print("ok")
