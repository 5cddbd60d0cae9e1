import torch, torch.nn.functional as F, sys
dev = sys.argv[1] if len(sys.argv) > 1 else "cpu"
torch.manual_seed(0)
for (n_in, kw) in [(24, dict(scale_factor=(1.205, 1.205))), (28, dict(scale_factor=(1.2, 1.2))), (33, dict(scale_factor=(1.2,1.2))), (39, dict(size=(48, 48))), (48, dict(scale_factor=(0.838, 0.838))), (40, dict(scale_factor=(0.798, 0.798))), (31, dict(size=(24, 24))),
                   (192, dict(scale_factor=(1.205, 1.205))), (384, dict(scale_factor=(0.838,0.838))), (321, dict(scale_factor=(0.798,0.798))), (256, dict(size=(192,192)))]:
    x = torch.randn(2, 3, n_in, n_in, dtype=torch.float64, device=dev, requires_grad=True)
    y = F.interpolate(x, **kw)
    v = torch.randn_like(y)
    (g,) = torch.autograd.grad(y, x, v)
    lhs = float((y * v).sum()); rhs = float((x * g).sum())
    # exact adjoint through the forward's own index map
    idx = F.interpolate(torch.arange(n_in, dtype=torch.float64, device=dev).view(1, 1, 1, n_in).expand(1, 1, n_in, n_in), **kw)[0, 0, 0].long()
    gt = torch.zeros_like(x)
    tmp = torch.zeros(2, 3, n_in, y.shape[3], dtype=torch.float64, device=dev).index_add_(2, idx, v)
    gt = torch.zeros(2, 3, n_in, n_in, dtype=torch.float64, device=dev).index_add_(3, idx, tmp)
    print(dev, n_in, kw, "out", y.shape[-1], "<Lx,v>-<x,LTv> rel %.2e" % (abs(lhs - rhs) / abs(lhs)), "|g - exact adjoint| rel %.3e" % float((g - gt).norm() / gt.norm()))
