"""Times the InstanceNorm2d+ReLU cluster kernels at the three trunk sites (per-GPU batch 8 and 16 at a 768^2 crop)
next to ATen's F.instance_norm + relu.  CUDA events, L2 flushed between launches.

    python tools/bench_instnorm.py            (MRFP_IN_SLICE_KB=72 default; try 144 / 200 in a fresh process)
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrfp_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def t(fn, iters=10):
        ts = []
        for i in range(iters + 3):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    rows = []
    for (n, c, hw) in [(8, 64, 384 * 384), (8, 256, 192 * 192), (8, 512, 96 * 96), (16, 256, 192 * 192)]:
        side = int(hw ** 0.5)
        x = torch.randn(n, c, side, side, device=dev) * 2 + 1
        gy = torch.randn_like(x)
        w = 1 + 0.2 * torch.randn(c, device=dev)
        b = 0.2 * torch.randn(c, device=dev)
        y = torch.empty_like(x); gx = torch.empty_like(x)
        mean = torch.empty(n, c, device=dev); inv = torch.empty(n, c, device=dev)
        dg = torch.empty(n, c, device=dev); db = torch.empty(n, c, device=dev)
        psum = torch.empty(n, c, device=dev, dtype=torch.float64)
        tf = t(lambda: _lib.check(lib.mrfp_instnorm_fwd_f32(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), mean.data_ptr(),
                                                            inv.data_ptr(), psum.data_ptr(), n, c, hw, 1e-5, 1, st), "fwd"))
        tb = t(lambda: _lib.check(lib.mrfp_instnorm_bwd_f32(gy.data_ptr(), x.data_ptr(), w.data_ptr(), b.data_ptr(), mean.data_ptr(),
                                                            inv.data_ptr(), gx.data_ptr(), dg.data_ptr(), db.data_ptr(), n, c, hw, 1, st), "bwd"))
        ta = t(lambda: F.relu(F.instance_norm(x, weight=w, bias=b)))
        xr = x.clone().requires_grad_(True)
        yr = F.relu(F.instance_norm(xr, weight=w, bias=b))
        tab = t(lambda: torch.autograd.grad(yr, xr, gy, retain_graph=True))
        nbytes = x.numel() * 4
        rows.append({"shape": [n, c, side, side], "fwd_us": tf * 1e3, "fwd_gbs": 2 * nbytes / tf / 1e6, "bwd_us": tb * 1e3,
                     "bwd_gbs": 3 * nbytes / tb / 1e6, "aten_fwd_us": ta * 1e3, "aten_bwd_us": tab * 1e3})
        del x, gy, y, gx, xr, yr
    print(json.dumps({"slice_kb": os.environ.get("MRFP_IN_SLICE_KB", "72"), "rows": rows}))


if __name__ == "__main__":
    main()
