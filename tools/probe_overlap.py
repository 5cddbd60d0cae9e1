"""Does an HBM-bound element-wise kernel overlap a tensor-bound tcgen05 conv launched on another stream?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrfp_b200 import _lib
lib = _lib.load()
fn = lib.mrfp_debug_conv3x3_bf16
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 4
n, hw, ci, co = 8, 384, 256, 128
a = torch.randn(n, hw, hw, ci, device="cuda").to(torch.bfloat16)
w = (torch.randn(9, co, ci, device="cuda") * 0.02).to(torch.bfloat16)
y = torch.empty(n, hw, hw, co, device="cuda", dtype=torch.bfloat16)
src = torch.randn(300 << 20, device="cuda", dtype=torch.bfloat16)      # 600 MB
dst = torch.empty_like(src)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def conv(stream):
    assert fn(a.data_ptr(), w.data_ptr(), y.data_ptr(), n, hw, hw, ci, co, 1, None, None, None, stream.cuda_stream) == 0

def elem(stream):
    with torch.cuda.stream(stream):
        torch.clamp_min(src, 0, out=dst)

def timed(f, reps=5):
    ts = []
    for _ in range(reps + 2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[2:])[len(ts[2:]) // 2] * 1e3

def both(order):
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    if order == "conv_first":
        conv(s1); elem(s2)
    else:
        elem(s2); conv(s1)

def only_conv():
    s1.wait_stream(torch.cuda.current_stream()); conv(s1)
def only_elem():
    s2.wait_stream(torch.cuda.current_stream()); elem(s2)

print("conv alone   %.1f us" % timed(only_conv))
print("relu 1.2GB alone %.1f us" % timed(only_elem))
print("conv first, both streams %.1f us" % timed(lambda: both("conv_first")))
print("elem first, both streams %.1f us" % timed(lambda: both("elem_first")))
