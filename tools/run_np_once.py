"""One NP+ forward + backward launch on a given shape after an L2 flush (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrfp_b200 import _lib

n, c, h, w = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "8,256,192,192").split(","))
lib = _lib.load()
x = torch.relu(torch.randn(n, c, h, w, device="cuda")); out = torch.empty_like(x)
alpha = 1 + 0.75 * torch.randn(n, c, device="cuda"); eps = 0.75 * torch.randn(n, c, device="cuda")
mean = torch.empty(n, c, device="cuda"); beta = torch.empty(n, c, device="cuda")
wsb = lib.mrfp_npplus_ws_bytes(n, c, h * w)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    flush.zero_()
    assert lib.mrfp_npplus_fwd_f32(x.data_ptr(), alpha.data_ptr(), eps.data_ptr(), out.data_ptr(), mean.data_ptr(),
                                   beta.data_ptr(), ws.data_ptr(), wsb, n, c, h * w, st) == 0
    flush.zero_()
    assert lib.mrfp_npplus_bwd_f32(x.data_ptr(), alpha.data_ptr(), eps.data_ptr(), mean.data_ptr(), out.data_ptr(),
                                   ws.data_ptr(), wsb, n, c, h * w, st) == 0
torch.cuda.synchronize()
