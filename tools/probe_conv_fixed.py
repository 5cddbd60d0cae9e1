"""How much of a small conv launch is fixed cost: the 64->64 @192^2 layer at batch 8, 16, 32, 64."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrfp_b200 import _lib
lib = _lib.load()
fn = lib.mrfp_debug_conv3x3_bf16
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 4
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (ci, co, hw) in ((64, 64, 192), (128, 64, 384)):
    for n in (1, 2, 4, 8, 16, 32):
        a = torch.randn(n, hw, hw, ci, device="cuda").to(torch.bfloat16)
        w = (torch.randn(9, co, ci, device="cuda") * 0.05).to(torch.bfloat16)
        y = torch.empty(n, hw, hw, co, device="cuda", dtype=torch.bfloat16)
        ts = []
        for i in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(a.data_ptr(), w.data_ptr(), y.data_ptr(), n, hw, hw, ci, co, 1, None, None, None, st); e1.record()
            torch.cuda.synchronize()
            if i >= 3: ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        tiles = n * ((hw + 15) // 16) * ((hw + 15) // 16)
        print(f"{ci}->{co} @{hw} N={n:2d}: {t*1e3:7.1f} us, tiles/SM {tiles/148:6.2f}, us per tile-wave {t*1e3/((tiles+147)//148):6.2f}, {2.0*n*hw*hw*co*9*ci/t/1e9:7.1f} TFLOP/s")
