"""TEMP: ablation of the tcgen05 conv on the chain's shapes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrfp_b200 import _lib

lib = _lib.load()
fn = lib.mrfp_debug_conv3x3_bf16
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 4
setdbg = lib.mrfp_debug_conv_set
setdbg.argtypes = [ctypes.c_int]
n = 8
shapes = [(64, 64, 1, 192), (64, 64, 2, 321), (64, 128, 2, 277), (128, 64, 1, 384), (64, 128, 1, 384), (128, 256, 2, 332), (256, 128, 1, 384)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
variants = [("base", 0, False), ("stats", 0, True), ("noepi", 2, False), ("nomma", 4, False), ("noA", 8, False), ("noB", 16, False),
            ("noAB", 24, False), ("noAB_noepi", 26, False), ("nomma_noepi", 6, False), ("noAB_nomma", 28, False)]
for (ci, co, dil, hw) in shapes:
    a = torch.randn(n, hw, hw, ci, device="cuda").to(torch.bfloat16)
    w = (torch.randn(9, co, ci, device="cuda") * (2.0 / (9 * ci)) ** 0.5).to(torch.bfloat16)
    y = torch.empty(n, hw, hw, co, device="cuda", dtype=torch.bfloat16)
    cnt = torch.ones(hw + 64, dtype=torch.int32, device="cuda"); cnt[hw:] = 0
    acc = torch.zeros(2 * 256, dtype=torch.float64, device="cuda")
    fl = 2.0 * n * hw * hw * co * 9 * ci
    line = f"{ci:3d}->{co:3d} d{dil} @{hw}:"
    for name, flag, stats in variants:
        setdbg(flag)
        ts = []
        for i in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(a.data_ptr(), w.data_ptr(), y.data_ptr(), n, hw, hw, ci, co, dil, cnt.data_ptr() if stats else None,
                    cnt.data_ptr() if stats else None, acc.data_ptr() if stats else None, st)
            e1.record(); torch.cuda.synchronize()
            assert rc == 0, rc
            if i >= 2:
                ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        line += f" {name} {t * 1e3:.0f}us({fl / t / 1e9:.0f})"
    setdbg(0)
    print(line, flush=True)
    del a, w, y
