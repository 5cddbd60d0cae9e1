"""Golden fixtures for the trunk's InstanceNorm2d(affine) + ReLU, from the REFERENCE's own layers.

    python tests/golden/make_golden_instnorm.py        (build container only: imports /root/reference)

Builds the reference's `Bottleneck(..., iw=4)` (network/Resnet.py:148-227), takes ITS `instance_norm_layer` /
`relu` objects, fills gamma / beta with seeded values and evaluates `relu(instance_norm_layer(x))` and its autograd
gradients on seeded inputs (tests.common.make_feat shifted to both signs).  Writes instnorm.npz next to this script.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference           # noqa: E402
from tests.common import make_in_case                # noqa: E402

load_reference()
from network import Resnet                           # noqa: E402  (the reference's module)

CASES = {"a": (2, 8, 12, 12), "b": (3, 4, 7, 9), "c": (2, 16, 33, 31), "d": (1, 4, 64, 72)}


def main():
    out = {}
    for i, (name, shape) in enumerate(CASES.items()):
        n, c, h, w = shape
        blk = Resnet.Bottleneck(c, c // 4, iw=4)
        x_np, gamma, beta, gy = make_in_case(400 + i, shape)
        with torch.no_grad():
            blk.instance_norm_layer.weight.copy_(torch.from_numpy(gamma))
            blk.instance_norm_layer.bias.copy_(torch.from_numpy(beta))
        x = torch.from_numpy(x_np).requires_grad_(True)
        y = blk.relu(blk.instance_norm_layer(x) * 1.0)      # (* 1.0: the reference's ReLU is in-place)
        y.backward(torch.from_numpy(gy))
        out[f"{name}_shape"] = np.array(shape)
        out[f"{name}_y"] = y.detach().numpy()
        out[f"{name}_gx"] = x.grad.numpy()
        out[f"{name}_gw"] = blk.instance_norm_layer.weight.grad.numpy()
        out[f"{name}_gb"] = blk.instance_norm_layer.bias.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "instnorm.npz"), **out)
    print("instnorm.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
