"""Gathered forward convolutions (csrc/conv_gather.cu) of the 7 fused stages at batch 8, alone: time without / with the BN
statistics in the epilogue, next to the tap kernel on the materialised operand, and parity between the two.

    python tools/bench_gather.py
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrfp_b200 import _lib

lib = _lib.load()
tap = lib.mrfp_debug_conv3x3_bf16
tap.restype = ctypes.c_int
tap.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 4
gat = lib.mrfp_debug_conv3x3_gather_fwd
gat.restype = ctypes.c_int
gat.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 5 + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 4

n = int(os.environ.get("MRFP_PROFILE_N", "8"))
# (cin, cout, dil, conv resolution, source resolution): stages 1..7 of deepv3.py:320-327 at a 768^2 crop
stages = [(64, 64, 1, 231, 192), (64, 128, 2, 277, 231), (128, 256, 2, 332, 277), (256, 128, 1, 384, 332), (128, 64, 1, 384, 384),
          (64, 64, 2, 321, 384), (64, 64, 2, 256, 321)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def timed(fn):
    ts = []
    for i in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn()
        e1.record(); torch.cuda.synchronize()
        assert rc == 0, rc
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


for k, (ci, co, dil, hw, shw) in enumerate(stages, start=1):
    torch.manual_seed(k)
    y = torch.randn(n, shw, shw, ci, device="cuda").to(torch.bfloat16)
    w = (torch.randn(9, co, ci, device="cuda") * (2.0 / (9 * ci)) ** 0.5).to(torch.bfloat16)
    stats = torch.zeros(4, 256, device="cuda")
    stats[2, :ci] = 0.5 + torch.rand(ci, device="cuda"); stats[3, :ci] = 0.2 * torch.randn(ci, device="cuda")
    scale = torch.tensor(shw / hw, dtype=torch.float32)
    idx = torch.clamp(torch.floor(torch.arange(hw, dtype=torch.float32) * scale).to(torch.int64), max=shw - 1).cuda()
    idx32 = idx.to(torch.int32).contiguous()
    out_g = torch.empty(n, hw, hw, co, device="cuda", dtype=torch.bfloat16)
    out_t = torch.empty_like(out_g)
    a = torch.relu(y.float().index_select(1, idx).index_select(2, idx) * stats[2, :ci] + stats[3, :ci]).to(torch.bfloat16).contiguous()
    cnt = torch.zeros(hw + 64, dtype=torch.int32, device="cuda"); cnt[:hw] = 1
    acc = torch.zeros(2 * 256, dtype=torch.float64, device="cuda")
    t_tap = timed(lambda: tap(a.data_ptr(), w.data_ptr(), out_t.data_ptr(), n, hw, hw, ci, co, dil, None, None, None, st))
    t_tap_s = timed(lambda: tap(a.data_ptr(), w.data_ptr(), out_t.data_ptr(), n, hw, hw, ci, co, dil, cnt.data_ptr(), cnt.data_ptr(),
                                acc.data_ptr(), st))
    line = f"stage {k}: {ci:3d}->{co:3d} d{dil} @{hw} <- {shw}: tap {t_tap:7.1f} / with stats {t_tap_s:7.1f} us"
    if True:
        t = timed(lambda: gat(y.data_ptr(), shw, shw, idx32.data_ptr(), idx32.data_ptr(), stats.data_ptr(), w.data_ptr(),
                              out_g.data_ptr(), n, hw, hw, ci, co, dil, None, None, None, st))
        acc.zero_()
        ts = timed(lambda: gat(y.data_ptr(), shw, shw, idx32.data_ptr(), idx32.data_ptr(), stats.data_ptr(), w.data_ptr(),
                               out_g.data_ptr(), n, hw, hw, ci, co, dil, cnt.data_ptr(), cnt.data_ptr(), acc.data_ptr(), st))
        err = float((out_g.float() - out_t.float()).norm() / out_t.float().norm())
        line += f" | gathered {t:7.1f} / {ts:7.1f} us (rel l2 vs tap {err:.2e})"
    print(line, flush=True)
    del y, w, a, out_g, out_t
