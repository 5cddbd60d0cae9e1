"""Per-kernel GPU time of one bench step (torch profiler / CUPTI), plus event-timed step time.

    python tools/profile_step.py [-v]        MRFP_PROFILE_N: per-GPU batch (default 8)
The step is bench.py's: chain fwd with NP+ call 1 folded in, NP+ call 2, tail through the classifier, backward of all."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from mrfp_b200 import hrfp as H, npplus as NP
from mrfp_b200.model import init_hrfp_module
dev = "cuda"
torch.manual_seed(1)
n = int(os.environ.get("MRFP_PROFILE_N", "8"))
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

def step():
    a = xp.detach().requires_grad_(True); b = f2.detach().requires_grad_(True); d = d1.detach().requires_grad_(True)
    final2.weight.grad = None; final2.bias.grad = None
    x, dec = H.hrfp_chain(a, convs, bns, 768, 768, np_draws=draws[0], math_mode=H.MATH_BF16, lazy_dec=True)
    y2 = NP.np_plus_with_draws(b, *draws[1])
    o = H.hrfp_plus_final2(d, final2, dec)
    torch.autograd.backward([x, o, y2], [g_x, g_d, g_f])

for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record(); torch.cuda.synchronize()
print("step %.3f ms (10 iterations, CUDA events)" % (e0.elapsed_time(e1) / 10))
iters = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
rows = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        rows.append((ev.time_range.start, ev.name, ev.time_range.end - ev.time_range.start))
rows.sort()
per = len(rows) // iters
last = rows[-per:]
agg = collections.OrderedDict()
for _, name, dur in last:
    short = name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:60]
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += dur
tot = sum(v[1] for v in agg.values())
print("kernels per step: %d, sum of kernel time %.1f us" % (per, tot))
for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s %3d %9.1f us %5.1f %%" % (k, c, d, 100 * d / tot))
if "-v" in sys.argv:
    for _, name, dur in last:
        print("   %-90s %8.1f" % (name[:90], dur))
