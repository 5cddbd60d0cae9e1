"""HRFP chain timing at BASELINE config-2 shapes (B=8, 768^2 crop): ours (bf16 tcgen05 / fp32) vs eager torch on the GPU."""
import math, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import mrfp_oracle as O
from tests.common import make_hrfp_params
from tests.test_hrfp_gpu import _modules
from mrfp_b200.hrfp import hrfp_chain

FLOP_FWD_PER_SAMPLE = 204.12e9    # dense count (what both implementations execute)

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

def eager_chain(convs, bns, xp, h, w):
    o = F.relu(bns[0](F.interpolate(convs[0](xp), scale_factor=(1.205, 1.205))))
    o = F.relu(bns[1](F.interpolate(convs[1](o), scale_factor=(1.2, 1.2))))
    o = F.relu(bns[2](F.interpolate(convs[2](o), scale_factor=(1.2, 1.2))))
    d = F.relu(bns[3](F.interpolate(convs[3](o), size=(h // 2, w // 2))))
    o = F.relu(bns[4](F.interpolate(convs[4](d), size=(h // 2, w // 2))))
    o = F.relu(bns[5](F.interpolate(convs[5](o), scale_factor=(0.838, 0.838))))
    o = F.relu(bns[6](F.interpolate(convs[6](o), scale_factor=(0.798, 0.798))))
    o = F.relu(bns[7](F.interpolate(convs[7](o), size=(math.ceil(h / 4), math.ceil(w / 4)))))
    return o, d

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    h = w = 768
    ws, gs = make_hrfp_params(1)
    convs, bns = _modules(ws, gs, "cuda")
    torch.manual_seed(0)
    xp = torch.relu(torch.randn(n, 64, 192, 192, device="cuda")).requires_grad_(True)
    g1 = torch.randn(n, 64, 192, 192, device="cuda"); g2 = torch.randn(n, 256, 384, 384, device="cuda")
    res = {"batch": n}
    for mode, name in ((2, "bf16_tc"), (0, "fp32_cc")):
        if mode == 0 and n > 2: continue
        def fwd():
            return hrfp_chain(xp, convs, bns, h, w, math_mode=mode)
        def fwdbwd():
            xp.grad = None
            o, d = fwd()
            torch.autograd.backward([o, d], [g1, g2])
        tf = timeit(fwd); tfb = timeit(fwdbwd)
        res[name] = dict(fwd_ms=tf, fwd_bwd_ms=tfb, fwd_tflops=FLOP_FWD_PER_SAMPLE * n / tf / 1e9,
                         fwd_bwd_tflops=2 * FLOP_FWD_PER_SAMPLE * n / tfb / 1e9)
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        for m in bns: m.train()
        def fwd():
            return eager_chain(convs, bns, xp, h, w)
        def fwdbwd():
            xp.grad = None
            o, d = fwd()
            torch.autograd.backward([o, d], [g1, g2])
        try:
            tf = timeit(fwd, 3, 1); tfb = timeit(fwdbwd, 3, 1)
            res["eager_tf32" if tf32 else "eager_fp32"] = dict(fwd_ms=tf, fwd_bwd_ms=tfb)
        except Exception as e:
            res["eager_tf32" if tf32 else "eager_fp32"] = str(e)[:100]
    print(json.dumps(res, indent=1))

if __name__ == "__main__":
    main()
