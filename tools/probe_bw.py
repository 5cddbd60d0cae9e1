import torch, sys
def t(fn, n=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(n + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2]
for shape in [(8, 256, 192, 192), (8, 64, 192, 192)]:
    x = torch.randn(*shape, device="cuda"); y = torch.empty_like(x)
    mb = x.numel() * 4 / 1e6
    ms = t(lambda: x.sum((2, 3)))
    print(shape, "plane-sum (read only): %.1f us  %.0f GB/s" % (ms * 1e3, mb / ms / 1e3))
    ms = t(lambda: torch.mul(x, 1.5, out=y))
    print(shape, "scale out-of-place (R+W): %.1f us  %.0f GB/s" % (ms * 1e3, 2 * mb / ms / 1e3))
    ms = t(lambda: y.copy_(x))
    print(shape, "copy (R+W): %.1f us  %.0f GB/s" % (ms * 1e3, 2 * mb / ms / 1e3))
    # no flush: second pass right after a read pass (L2 warm for the tail)
    def two():
        x.sum((2, 3)); torch.mul(x, 1.5, out=y)
    ms = t(two)
    print(shape, "sum then scale back-to-back: %.1f us -> algorithmic %.0f GB/s" % (ms * 1e3, 2 * mb / ms / 1e3))
