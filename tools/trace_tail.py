"""Phase timeline of the classifier-tail kernels: clock64 sums of thread 0 of every CTA per phase of the tile loop.

    MRFP_EXTRA_NVCC_FLAGS=-DMRFP_TAIL_TRACE python -m mrfp_b200.build --force && python tools/trace_tail.py
The instrumentation is compiled out of the default build (rebuild without the flag afterwards)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrfp_b200 import hrfp as H, _lib
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
d1 = torch.randn(n, 256, 192, 192, device=dev)
g_d = torch.randn(n, 19, 384, 384, device=dev)
lib = _lib.load()
lib = ctypes.CDLL(_lib.LIB_PATH)
lib.mrfp_debug_tail_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
def step():
    a = xp.detach().requires_grad_(True); d = d1.detach().requires_grad_(True)
    x, dec = H.hrfp_chain(a, convs, bns, 768, 768, math_mode=H.MATH_BF16, lazy_dec=True, want_out=False)
    o = H.hrfp_plus_final2(d, final2, dec)
    o.backward(g_d)
for _ in range(3):
    step()
torch.cuda.synchronize()
lib.mrfp_debug_tail_trace(None, 1)
step(); torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 16)()
lib.mrfp_debug_tail_trace(buf, 0)
ctas, tiles = 296, 8 * 384 * 6
names_f = ["wait A", "issue+pix", "mbar wait", "-", "sync B", "blend+contract", "sync C", "epilogue"]
names_b = ["wait A", "issue", "spx", "mbar wait", "gC, sync B", "(1) dA", "(2) gW2", "-"]
for off, names in ((0, names_f), (8, names_b)):
    tot = sum(buf[off + k] for k in range(8))
    print("fwd2" if off == 0 else "bwd2", "cycles per CTA %.0f  (= %.1f us at 1.9 GHz), per tile %.0f" % (tot / ctas, tot / ctas / 1900, tot / tiles))
    for k in range(8):
        print("   %-18s %8.0f cyc/tile  %5.1f %%" % (names[k], buf[off + k] / tiles, 100.0 * buf[off + k] / max(tot, 1)))
