"""Times the tcgen05 conv kernel on the chain's forward and dgrad shapes (B=8, 768^2 crop) (the tap kernel conv3x3_tc_kernel through its debug hook)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrfp_b200 import _lib

lib = _lib.load()
fn = lib.mrfp_debug_conv3x3_bf16
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 4
n = 8
stages = [(64, 64, 1, 192), (64, 64, 1, 231), (64, 128, 2, 277), (128, 256, 2, 332), (256, 128, 1, 384), (128, 64, 1, 384),
          (64, 64, 2, 321), (64, 64, 2, 256)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
tot = {"fwd": 0.0, "bwd": 0.0}
for direction in ("fwd", "bwd"):
    for k, (cin, cout, dil, hw) in enumerate(stages):
        ci, co = (cin, cout) if direction == "fwd" else (cout, cin)
        a = torch.randn(n, hw, hw, ci, device="cuda").to(torch.bfloat16)
        w = (torch.randn(9, co, ci, device="cuda") * (2.0 / (9 * ci)) ** 0.5).to(torch.bfloat16)
        y = torch.empty(n, hw, hw, co, device="cuda", dtype=torch.bfloat16)
        ts = []
        for i in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(a.data_ptr(), w.data_ptr(), y.data_ptr(), n, hw, hw, ci, co, dil, None, None, None, st)
            e1.record(); torch.cuda.synchronize()
            assert rc == 0, rc
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        fl = 2.0 * n * hw * hw * co * 9 * ci
        tot[direction] += t
        print(f"{direction} stage {k}: {ci:3d}->{co:3d} d{dil} @{hw}: {t * 1e3:7.1f} us  {fl / t / 1e9:7.1f} TFLOP/s")
        del a, w, y
print("total fwd %.1f us, dgrad %.1f us" % (tot["fwd"] * 1e3, tot["bwd"] * 1e3))
