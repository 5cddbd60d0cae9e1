"""Micro-benchmark of the NP+ kernels (CUDA events, L2 flushed between iterations)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrfp_b200 import _lib


def bench(n, c, h, w, iters=20):
    lib = _lib.load()
    x = torch.relu(torch.randn(n, c, h, w, device="cuda"))
    out = torch.empty_like(x)
    alpha = 1 + 0.75 * torch.randn(n, c, device="cuda")
    eps = 0.75 * torch.randn(n, c, device="cuda")
    mean = torch.empty(n, c, device="cuda")
    beta = torch.empty(n, c, device="cuda")
    wsb = lib.mrfp_npplus_ws_bytes(n, c, h * w)
    ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for name in ("fwd", "bwd"):
        ts = []
        for i in range(iters + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if name == "fwd":
                rc = lib.mrfp_npplus_fwd_f32(x.data_ptr(), alpha.data_ptr(), eps.data_ptr(), out.data_ptr(),
                                             mean.data_ptr(), beta.data_ptr(), ws.data_ptr(), wsb, n, c, h * w, st)
            else:
                rc = lib.mrfp_npplus_bwd_f32(x.data_ptr(), alpha.data_ptr(), eps.data_ptr(), mean.data_ptr(),
                                             out.data_ptr(), ws.data_ptr(), wsb, n, c, h * w, st)
            e1.record()
            assert rc == 0, rc
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        gb = 2 * x.numel() * 4 / 1e9
        res[name] = dict(ms=med, min_ms=ts[0], gbps=gb / (med * 1e-3), gbps_best=gb / (ts[0] * 1e-3))
    # eager reference-style implementation on the same GPU for context
    def eager(feat):
        m = feat.mean((2, 3), keepdim=True)
        d = torch.std(m, 0, keepdim=True)
        s = d / d.max() * 1.5
        a = alpha.view(n, c, 1, 1); b = 1 + eps.view(n, c, 1, 1) * s
        return a * feat - a * m + b * m
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = eager(x); e1.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1))
    ts.sort()
    res["eager_fwd_ms"] = ts[len(ts) // 2]
    return res


if __name__ == "__main__":
    if len(sys.argv) > 1:      # e.g. 8,256,192,192
        for a in sys.argv[1:]:
            shape = tuple(int(v) for v in a.split(","))
            r = bench(*shape)
            print(shape, "fwd %.1f us %.0f GB/s | bwd %.1f us %.0f GB/s" % (r["fwd"]["ms"] * 1e3, r["fwd"]["gbps"], r["bwd"]["ms"] * 1e3, r["bwd"]["gbps"]))
        sys.exit(0)
    for shape in [(8, 64, 192, 192), (8, 256, 192, 192), (2, 64, 192, 192), (2, 256, 192, 192),
                  (8, 16, 384, 384), (8, 116, 96, 96)]:
        print(shape, json.dumps(bench(*shape)))
