"""Per-phase timing of the NP+ ring kernel from in-kernel %globaltimer stamps (MRFP_NPPLUS_TRACE=1)."""
import os, sys
os.environ["MRFP_NPPLUS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mrfp_b200 import _lib

def run(n, c, h, w, reps=6):
    lib = _lib.load()
    x = torch.relu(torch.randn(n, c, h, w, device="cuda")); out = torch.empty_like(x)
    alpha = 1 + 0.75 * torch.randn(n, c, device="cuda"); eps = 0.75 * torch.randn(n, c, device="cuda")
    mean = torch.empty(n, c, device="cuda"); beta = torch.empty(n, c, device="cuda")
    base = (lib.mrfp_npplus_ws_bytes(n, c, h * w) + 7) // 8 * 8
    wsb = base + 148 * 64
    ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for r in range(reps):
        flush.zero_()
        rc = lib.mrfp_npplus_fwd_f32(x.data_ptr(), alpha.data_ptr(), eps.data_ptr(), out.data_ptr(), mean.data_ptr(),
                                     beta.data_ptr(), ws.data_ptr(), wsb, n, c, h * w, st)
        assert rc == 0
        torch.cuda.synchronize()
    t = ws[base:].cpu().numpy().view(np.uint64).reshape(148, 8)[:, :8].astype(np.int64)
    t0 = t[:, 0].min()
    names = ["start", "phaseA_end", "after_gridsync", "after_stats", "end", "after_stage1", "after_sync2", "after_reduce"]
    print((n, c, h, w))
    for i, nm in enumerate(names):
        col = (t[:, i] - t0) / 1e3
        print(f"  {nm:15s} min {col.min():8.2f}  median {np.median(col):8.2f}  max {col.max():8.2f} us")

if __name__ == "__main__":
    for shp in [(8, 64, 192, 192), (8, 256, 192, 192), (2, 64, 192, 192)]:
        run(*shp)
