"""Import shim for the read-only reference at /root/reference (build container only).

TEST INFRASTRUCTURE ONLY — used by tests/golden/make_golden.py and by container-only tests to run
the reference's own code (deepv3.MRFPPlus) as the ground truth.  /root/reference does not exist on
the GPU box, so nothing that runs there may call `load_reference()`.

Stubs the packages the reference imports but never uses on the MRFP path (deepv3.py:36-37, 48-58,
network/cov_settings.py:4), blocks the pretrained-weight download (network/Resnet.py:659) and forces
plain BatchNorm2d (config.py:93 defaults to SyncBatchNorm).
"""
import os
import sys
import types

REF_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "deepv3.py"))


def load_reference():
    """Returns the reference `deepv3` module (with `MRFPPlus`)."""
    import torch
    if "deepv3" in sys.modules and getattr(sys.modules["deepv3"], "_mrfp_shimmed", False):
        return sys.modules["deepv3"]
    sys.dont_write_bytecode = True            # /root/reference is read-only
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    stub("matplotlib"); stub("matplotlib.pyplot")
    stub("pytorch_wavelets", DWTForward=_Dummy, DWTInverse=_Dummy)
    stub("segmentation_models_pytorch")
    stub("segmentation_models_pytorch.base", SegmentationModel=_Dummy, SegmentationHead=_Dummy,
         ClassificationHead=_Dummy, modules=types.ModuleType("modules"))
    stub("segmentation_models_pytorch.decoders")
    stub("segmentation_models_pytorch.decoders.unet", UnetDecoder=_Dummy)
    stub("segmentation_models_pytorch.encoders", get_encoder=lambda *a, **k: None)
    stub("kmeans1d")
    import torch.utils.model_zoo as mz
    mz.load_url = lambda *a, **k: {}
    import config
    config.cfg.MODEL.BNFUNC = torch.nn.BatchNorm2d
    import deepv3
    deepv3._mrfp_shimmed = True
    return deepv3
