"""Gather-form backward of the reference's Upsample (network/mynn.py:114-119; csrc/bilinear.cu) vs ATen's scatter."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [((2, 19, 96, 96), (192, 192)), ((2, 256, 24, 20), (48, 40)), ((1, 3, 7, 33), (21, 40)),
                                  ((2, 5, 12, 12), (12, 12)), ((1, 19, 192, 192), (768, 768)), ((3, 7, 10, 9), (48, 31)),
                                  ((1, 2, 1, 1), (5, 4))])
def test_gather_backward_equals_aten_backward(case):
    from mrfp_b200.bilinear import bilinear_up_backward, upsample_bilinear
    shape, size = case
    torch.manual_seed(3)
    x = torch.randn(*shape, device="cuda", requires_grad=True)
    g = torch.randn(shape[0], shape[1], *size, device="cuda")
    ref = torch.ops.aten.upsample_bilinear2d_backward(g, list(size), list(shape), True, None, None)
    got = bilinear_up_backward(g, shape[2:])
    assert got.shape == ref.shape
    # fp64 adjoint: the interpolation is linear, so autograd of the fp64 forward is the exact transpose
    x64 = torch.zeros(*shape, device="cuda", dtype=torch.float64, requires_grad=True)
    (g64,) = torch.autograd.grad(F.interpolate(x64, size=size, mode="bilinear", align_corners=True), x64, g.double())
    scale = float(g64.abs().max())
    # ATen's fp32 forward takes its interpolation weights from float arithmetic (src = o * float((L-1)/(O-1))): they sit
    # up to ~1e-7 * o from the exact ones.  The gather uses the same float weights, so it is the adjoint of the fp32
    # forward and equals ATen's own backward up to summation order, and both sit ~1e-5 from the fp64 transpose.
    assert float((got.double() - ref.double()).abs().max()) <= 2e-6 * scale
    assert float((got.double() - g64).abs().max()) <= 1e-4 * scale
    # through autograd: module-level Upsample with the gather backward
    y = upsample_bilinear(x, size)
    assert torch.equal(y, F.interpolate(x.detach(), size=size, mode="bilinear", align_corners=True))
    y.backward(g)
    assert float((x.grad.double() - ref.double()).abs().max()) <= 2e-6 * scale


def test_bad_arguments():
    from mrfp_b200 import _lib
    lib = _lib.load()
    assert lib.mrfp_bilinear_bwd_table_bytes(8, 4) == 0
    assert lib.mrfp_bilinear_bwd_write_table(4, 8, None, 0) == -1
    assert lib.mrfp_bilinear_up_bwd_f32(None, None, 1, 1, 1, 1, 1, None, None, 1, None) == -1
