"""A/B of two builds / switches of the classifier tail at the benchmarked geometry (batch 2, 768^2).

    python tools/ab_tail.py dump /tmp/t1.pt          (build A)
    python tools/ab_tail.py dump /tmp/t2.pt          (build B)
    python tools/ab_tail.py cmp /tmp/t1.pt /tmp/t2.pt
`dump` also checks dec2 and the three classifier-side gradients against fp64 built from the materialised OCout_dec."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def dump(path):
    from mrfp_b200 import hrfp as H
    from mrfp_b200.model import init_hrfp_module
    dev = "cuda"
    torch.manual_seed(5)
    n, h, w = 2, 768, 768
    chans, dils = [64, 64, 64, 128, 256, 128, 64, 64, 64], [1, 1, 2, 2, 1, 1, 2, 2]
    convs = [torch.nn.Conv2d(chans[k], chans[k + 1], 3, padding=dils[k], dilation=dils[k]).to(dev).requires_grad_(False) for k in range(8)]
    bns = [torch.nn.BatchNorm2d(chans[k + 1]).to(dev).requires_grad_(False) for k in range(8)]
    for c, b in zip(convs, bns):
        init_hrfp_module(c); init_hrfp_module(b)
    final2 = torch.nn.Conv2d(256, 19, 1).to(dev)
    xp = torch.relu(torch.randn(n, 64, 192, 192, device=dev))
    d1 = torch.randn(n, 256, 192, 192, device=dev)
    g = torch.randn(n, 19, 384, 384, device=dev)
    g_x = torch.randn(n, 64, 192, 192, device=dev)
    xa = xp.clone().requires_grad_(True); da = d1.clone().requires_grad_(True)
    full = os.environ.get("AB_TAIL_ENCODER_ONLY", "0") == "0"      # full chain: the tail's gradient joins in the stage-4 dgrad
    x, dec = H.hrfp_chain(xa, convs, bns, h, w, want_out=full, math_mode=H.MATH_BF16, lazy_dec=True, update_running_stats=False)
    out = H.hrfp_plus_final2(da, final2, dec)
    if full:
        torch.autograd.backward([x, out], [g_x, g])
    else:
        out.backward(g)
    res = dict(out=out.detach(), g_d1=da.grad, g_w2=final2.weight.grad.clone(), g_b2=final2.bias.grad.clone(), g_xp=xa.grad)
    # fp64 from the materialised decoder feature of the same chain
    _, dec = H.hrfp_chain(xp, convs, bns, h, w, want_out=False, math_mode=H.MATH_BF16, lazy_dec=True, update_running_stats=False)
    full = H.hrfp_plus_add_upsampled(d1, dec).double().requires_grad_(True)
    w64 = final2.weight.detach().double().requires_grad_(True); b64 = final2.bias.detach().double().requires_grad_(True)
    ref = torch.nn.functional.conv2d(full, w64, b64)
    ref.backward(g.double())
    rel = lambda a, b: float((a.double() - b).norm() / b.norm())
    print("vs fp64: dec2 %.3e  g_W2 %.3e  g_b2 %.3e" % (rel(res["out"], ref.detach()), rel(res["g_w2"], w64.grad), rel(res["g_b2"], b64.grad)))
    torch.save({k: v.cpu() for k, v in res.items()}, path)


def cmp(p1, p2):
    a, b = torch.load(p1), torch.load(p2)
    for k in a:
        d = (a[k].double() - b[k].double())
        print("%-6s rel L2 %.3e  max abs %.3e  (max |ref| %.3e)" % (k, float(d.norm() / b[k].double().norm()), float(d.abs().max()), float(b[k].abs().max())))


if __name__ == "__main__":
    dump(sys.argv[2]) if sys.argv[1] == "dump" else cmp(sys.argv[2], sys.argv[3])
