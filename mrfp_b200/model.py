"""Drop-in `MRFPPlus` — the reference's torch.nn.Module surface for the MRFP hot path.

Mirrors /root/reference/deepv3.py:152-367 (`MRFPPlus`): same constructor arguments, same child-module
names and state_dict keys (the published MRFP+ checkpoint, README.md:18, loads with `strict=True`),
same `forward(x, gts=None, training=True)`, same `Normalization_Perturbation_Plus(feat)`, same three
`random.random()` gates and the same RNG consumption order of the HRFP re-randomisation
(deepv3.py:290-306 -> network/mynn.py:57-74).  The three MRFP insertion points call the sm_100a
kernels of libmrfp_b200.so (NP+: npplus.py; HRFP / HRFP+: hrfp.py); the surrounding DeepLabV3+ /
ResNet-50 host (stem, layer1-4, ASPP, decoder, loss) is plain PyTorch and is NOT accelerated here
(out of scope).

Differences from the reference, all observable only through the BN buffers of the 8 OC* BatchNorms:
the reference evaluates the HRFP chain on every forward, even when its result is discarded (eval mode,
or p >= 0.5 and p3 >= 0.5).  This module skips that dead compute; `strict_buffers=True` restores the
reference's buffer updates (running_mean / running_var / num_batches_tracked) in training mode.
"""
import math
import random

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import bilinear as _bilinear
from . import hrfp as _hrfp
from . import instnorm as _instnorm
from . import npplus as _npplus

# the trunk's InstanceNorm2d + ReLU pairs next to the insertion points run on the on-chip plane kernels
# (same result, forward 1R+1W); MRFP_FUSE_INSTNORM=0: ATen
FUSE_INSTNORM = os.environ.get("MRFP_FUSE_INSTNORM", "1") != "0"

HRFP_CONVS = ("OClayer1", "OClayer2", "OClayer3", "OClayer4",
              "OCdeclayer1", "OCdeclayer2", "OCdeclayer3", "OCdeclayer4")
HRFP_BNS = ("OC1_bn", "OC2_bn", "OC3_bn", "OC4_bn", "OC1_decbn", "OC2_decbn", "OC3_decbn", "OC4_decbn")


def upsample_bilinear(x, size):
    """network/mynn.py:114-119 (CUDA fp32: ATen forward, gather-form backward from csrc/bilinear.cu)."""
    return _bilinear.upsample_bilinear(x, size)


def init_hrfp_module(module: nn.Module):
    """network/mynn.py:57-74 for one conv or one BN (same torch RNG calls in the same order)."""
    if isinstance(module, nn.Conv2d):
        nn.init.kaiming_normal_(module.weight, nonlinearity="relu")
        if module.bias is not None:
            module.bias.data.zero_()
    elif isinstance(module, nn.BatchNorm2d):
        nn.init.normal_(module.weight, mean=0.0, std=0.5)
        module.bias.data.zero_()


def init_head(*models):
    """network/mynn.py:37-55 (`initialize_weights`)."""
    for model in models:
        for m in model.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()


class MRFPMixin:
    """The MRFP layers and insertion points, independent of the trunk (used by MRFPPlus below and usable
    on other trunks: SURVEY.md §8f-2)."""
    # fold NP+ call 1 into the HRFP chain when both gates are on (same result, two passes fewer); MRFP_FUSE_STEM_NP=0: off
    fuse_stem_np = os.environ.get("MRFP_FUSE_STEM_NP", "1") != "0"
    # take the NP+ statistics of call 2 in layer1's last ReLU (same result, NP+ forward becomes 1R+1W); MRFP_FUSE_LAYER1_NP=0: off
    fuse_layer1_np = os.environ.get("MRFP_FUSE_LAYER1_NP", "1") != "0"
    # HRFP+ tail: bilinear Upsample of dec1 evaluated inside the add kernel (no (N,256,h/2,w/2) intermediate); MRFP_FUSE_PLUS_TAIL=0: off
    fuse_plus_tail = os.environ.get("MRFP_FUSE_PLUS_TAIL", "1") != "0"
    # ... and through the classifier: final2(Upsample(dec1) + OCout_dec) in one kernel per direction; MRFP_FUSE_FINAL2=0: off
    fuse_final2 = os.environ.get("MRFP_FUSE_FINAL2", "1") != "0"

    def _build_hrfp(self, in_ch=64, widths=(64, 64, 128, 256)):
        chans = [in_ch, widths[0], widths[1], widths[2], widths[3], widths[2], widths[1], widths[0], in_ch]
        dils = [1, 1, 2, 2, 1, 1, 2, 2]
        for k, (cname, bname) in enumerate(zip(HRFP_CONVS, HRFP_BNS)):      # deepv3.py:221-237
            conv = nn.Conv2d(chans[k], chans[k + 1], kernel_size=3, stride=1, padding=dils[k], dilation=dils[k])
            setattr(self, cname, conv.requires_grad_(False))
            setattr(self, bname, nn.BatchNorm2d(chans[k + 1]).requires_grad_(False))
        self.reinit_hrfp()                                                  # deepv3.py:239-254

    def hrfp_modules(self):
        return [getattr(self, n) for n in HRFP_CONVS], [getattr(self, n) for n in HRFP_BNS]

    def reinit_hrfp(self):
        """deepv3.py:290-306: conv1, bn1, conv2, bn2, ... in module order."""
        for cname, bname in zip(HRFP_CONVS, HRFP_BNS):
            init_hrfp_module(getattr(self, cname))
            init_hrfp_module(getattr(self, bname))

    def Normalization_Perturbation_Plus(self, feat):
        """deepv3.py:268-277 on the fused NP+ kernels."""
        return _npplus.normalization_perturbation_plus(feat)

    def _plus_add(self, dec1_up, ocout_dec):
        """Insertion point 3 (deepv3.py:357)."""
        return _hrfp.hrfp_plus_add(dec1_up, ocout_dec)

    def mrfp_stem(self, xp, h, w, training, p, p2, p3):
        """Insertion point 1 (deepv3.py:316-330).  Returns (x, OCout_dec or None)."""
        x = xp
        want_out = training and p < 0.5
        want_dec = training and p3 < 0.5
        np_draws = None
        if training and p2 < 0.5:
            if want_out and self.fuse_stem_np:       # OCout + NP+(xp): NP+ rides on the chain's passes
                np_draws = _npplus.draw_np_plus_factors(xp)                 # same RNG position as the unfused call
            else:
                x = self.Normalization_Perturbation_Plus(xp)
        dec = None
        if want_out or want_dec:
            convs, bns = self.hrfp_modules()
            out, dec = _hrfp.hrfp_chain(xp, convs, bns, h, w, x_add=x if (want_out and np_draws is None) else None,
                                        want_out=want_out, want_dec=want_dec, math_mode=self.math_mode, lazy_dec=True,
                                        np_draws=np_draws)
            if want_out:
                x = out                                                     # OCout + x  (deepv3.py:330)
        elif training and self.strict_buffers:
            convs, bns = self.hrfp_modules()
            with torch.no_grad():
                _hrfp.hrfp_chain(xp, convs, bns, h, w, want_out=True, want_dec=False, math_mode=self.math_mode)
        return x, dec


class Bottleneck(nn.Module):
    """ResNet bottleneck with the optional trailing InstanceNorm of the reference's iw == 4 variant
    (network/Resnet.py:148-227); plain tensors instead of the reference's [x, w_arr] tuples."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, instance_norm=False):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample
        if instance_norm:
            self.instance_norm_layer = nn.InstanceNorm2d(planes * 4, affine=True)
        self.has_in = instance_norm
        self.emit_plane_sums = False       # set on the last block of layer1 by MRFPPlus: NP+ call 2 joins this block's IN + ReLU
        self.plane_sums = None
        self.np_applied = False
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        out = out + (x if self.downsample is None else self.downsample(x))
        if self.has_in and FUSE_INSTNORM and out.is_cuda and out.dtype == torch.float32:
            # IN + ReLU in one on-chip pass per plane; for layer1's last block also the plane sums NP+ call 2 needs
            if self.emit_plane_sums and self.training and out.shape[1] <= 256:
                # NP+ call 2 (deepv3.py:334-335) as one autograd node with its producer: the two draws are made
                # here, at the RNG position of the reference's call (nothing between this point and it consumes random numbers)
                alpha, eps = _npplus.draw_np_plus_factors(out)
                self.np_applied = True
                return _instnorm.module_instance_norm_relu_np_plus(self.instance_norm_layer, out, alpha, eps)
            if self.emit_plane_sums and self.training:
                out, self.plane_sums = _instnorm.module_instance_norm_relu(self.instance_norm_layer, out, True, True)
                return out
            self.plane_sums = None
            return _instnorm.module_instance_norm_relu(self.instance_norm_layer, out, True)
        if self.has_in:
            out = self.instance_norm_layer(out)
        if self.emit_plane_sums and out.is_cuda and self.training:
            # the block's ReLU also leaves the plane sums of its output for the NP+ call that follows
            out, self.plane_sums = _npplus.relu_with_plane_sums(out)
            return out
        self.plane_sums = None
        return self.relu(out)


def _make_layer(inplanes, planes, blocks, stride, in_last):
    downsample = None
    if stride != 1 or inplanes != planes * 4:
        downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False),
                                   nn.BatchNorm2d(planes * 4))
    layers = [Bottleneck(inplanes, planes, stride, downsample)]
    for i in range(1, blocks):
        layers.append(Bottleneck(planes * 4, planes, instance_norm=in_last and i == blocks - 1))
    return nn.Sequential(*layers)


class ASPP(nn.Module):
    """deepv3.py:64-126 (output stride 16: rates 6, 12, 18)."""

    def __init__(self, in_dim, reduction_dim=256, rates=(6, 12, 18)):
        super().__init__()
        feats = [nn.Sequential(nn.Conv2d(in_dim, reduction_dim, 1, bias=False), nn.BatchNorm2d(reduction_dim),
                               nn.ReLU(inplace=True))]
        for r in rates:
            feats.append(nn.Sequential(nn.Conv2d(in_dim, reduction_dim, 3, dilation=r, padding=r, bias=False),
                                       nn.BatchNorm2d(reduction_dim), nn.ReLU(inplace=True)))
        self.features = nn.ModuleList(feats)
        self.img_pooling = nn.AdaptiveAvgPool2d(1)
        self.img_conv = nn.Sequential(nn.Conv2d(in_dim, 256, 1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True))

    def forward(self, x):
        img = upsample_bilinear(self.img_conv(self.img_pooling(x)), x.shape[2:])
        return torch.cat([img] + [f(x) for f in self.features], 1)


class _ShuffleLayer0(nn.Module):
    """network/deepv3.py:121-151 with iw == 0: conv1 (conv-bn-relu) + maxpool under the reference's attribute name."""

    def __init__(self, conv1, maxpool):
        super().__init__()
        self.layer = nn.Sequential(conv1, maxpool)

    def forward(self, x):
        return self.layer(x)


class _ShuffleLayer4(nn.Module):
    """network/deepv3.py:153-179 with iw == 0: conv5 (conv-bn-relu)."""

    def __init__(self, conv5):
        super().__init__()
        self.layer = conv5

    def forward(self, x):
        return self.layer(x)


MOBILE_TRUNKS = ("mobilenetv2", "shufflenetv2")


class MRFPPlus(nn.Module, MRFPMixin):
    """DeepLabV3+ / ResNet-50 (IN at the stem and at the end of layer1, layer2: wt_layer=[0,0,4,4,4,0,0]) with
    MRFP+ applied while training.

    Extensions of this repo (the reference's MRFPPlus raises for any trunk but resnet-50,
    deepv3.py:177-178) — the same three insertion points on the other trunks of the reference's DeepV3Plus:
      trunk="resnet-101" (BASELINE config[3]): deep-stem ResNet-101, HRFP on its 128 stem channels;
      trunk="mobilenetv2" / "shufflenetv2" (BASELINE config[4]; network/deepv3.py:121-193, :259-283, forward :478-556,
      wt_layer all zero as DeepV3Plus forces): hooks after layer0 (16 ch @ stride 2 / 24 ch @ stride 4), after layer1
      (32 / 116 ch @ stride 8) and after final1.  The HRFP chain is sized RELATIVE TO xp (x1.205, x1.2, x1.2, 2x, 2x,
      x0.838, x0.798, 1x of the xp size — what deepv3.py:320-327 amounts to on a stride-4 stem), so a stride-2 stem
      peaks at the image size; its narrow stem is zero-padded to the kernels' channel granule inside the plan."""

    def __init__(self, num_classes, trunk="resnet-50", criterion=None, criterion_aux=None, variant="D16",
                 wt_layer=(0, 0, 4, 4, 4, 0, 0), use_wtloss=False, math_mode=_hrfp.MATH_BF16, strict_buffers=False):
        super().__init__()
        if trunk not in ("resnet-50", "resnet-101") + MOBILE_TRUNKS:
            raise ValueError("Not a valid network arch")                    # deepv3.py:177-178
        if trunk in MOBILE_TRUNKS:
            wt_layer = (0, 0, 0, 0, 0, 0, 0)                                # network/deepv3.py:120
        elif tuple(wt_layer) != (0, 0, 4, 4, 4, 0, 0):
            raise ValueError("only the reference's wt_layer=[0,0,4,4,4,0,0] host is provided")
        self.criterion, self.criterion_aux = criterion, criterion_aux
        self.variant, self.wt_layer, self.use_wtloss, self.trunk = variant, list(wt_layer), use_wtloss, trunk
        self.math_mode, self.strict_buffers = math_mode, strict_buffers
        self.hrfp_scale = 1                  # image size the chain is planned for = hrfp_scale x the real one (4 / stem stride)
        if trunk in MOBILE_TRUNKS:
            self._build_mobile(num_classes, trunk, variant)
            return

        if trunk == "resnet-50":                                            # Resnet.py:519-560 (7x7 stem), [3,4,6,3]
            stem_ch, blocks = 64, (3, 4, 6, 3)
            self.layer0 = nn.Sequential(nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False),
                                        nn.InstanceNorm2d(64, affine=True), nn.ReLU(inplace=True),
                                        nn.MaxPool2d(3, stride=2, padding=1))
        else:
            # SURVEY.md 8f-2 (extension: the reference's MRFPPlus only builds resnet-50): the reference's ResNet-101 is the
            # deep-stem ResNet3X3 (Resnet.py:352-433, :678-693), 128 channels at stride 4, [3,4,23,3]; wt_layer[2] == 4
            # puts the InstanceNorm on the third stem conv (layer0 laid out as network/deepv3.py:462-471 does)
            stem_ch, blocks = 128, (3, 4, 23, 3)
            self.layer0 = nn.Sequential(nn.Conv2d(3, 64, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                                        nn.Conv2d(64, 64, 3, stride=1, padding=1, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                                        nn.Conv2d(64, 128, 3, stride=1, padding=1, bias=False),
                                        nn.InstanceNorm2d(128, affine=True), nn.ReLU(inplace=True),
                                        nn.MaxPool2d(3, stride=2, padding=1))
        self.layer1 = _make_layer(stem_ch, 64, blocks[0], 1, True)
        self.layer2 = _make_layer(256, 128, blocks[1], 2, True)
        self.layer3 = _make_layer(512, 256, blocks[2], 2, False)
        self.layer4 = _make_layer(1024, 512, blocks[3], 2, False)
        for m in self.modules():                                            # network/Resnet.py:563-570
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        if variant == "D16":                                                # deepv3.py:184-189
            for n, m in self.layer4.named_modules():
                if "conv2" in n:
                    m.dilation, m.padding, m.stride = (2, 2), (2, 2), (1, 1)
                elif "downsample.0" in n:
                    m.stride = (1, 1)
        self.output_stride = 16
        self.aspp = ASPP(2048, 256)
        self.bot_fine = nn.Sequential(nn.Conv2d(256, 48, 1, bias=False), nn.BatchNorm2d(48), nn.ReLU(inplace=True))
        self.bot_aspp = nn.Sequential(nn.Conv2d(1280, 256, 1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True))
        self.final1 = nn.Sequential(nn.Conv2d(304, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                                    nn.Conv2d(256, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True))
        self.final2 = nn.Sequential(nn.Conv2d(256, num_classes, 1, bias=True))
        self._build_hrfp(stem_ch)                                           # HRFP runs on the stem's channel count
        init_head(self.aspp, self.bot_aspp, self.bot_fine, self.final1, self.final2)
        self.eps = 1e-5
        self.whitening = False
        self.three_input_layer = False

    def _build_mobile(self, num_classes, trunk, variant):
        """network/deepv3.py:121-193 (ShuffleNetV2 x1.0) and :259-283 (MobileNetV2): torchvision's modules — the
        reference's network/Mobilenet.py and Shufflenet.py are torchvision's files plus the (here unused) iw options —
        sliced and named as the reference slices them, so its DeepV3Plus state_dict (minus the `dsn` head) loads."""
        import torchvision
        if trunk == "mobilenetv2":
            f = torchvision.models.mobilenet_v2(weights=None).features
            self.layer0 = nn.Sequential(f[0], f[1])
            self.layer1 = nn.Sequential(*[f[i] for i in range(2, 7)])
            self.layer2 = nn.Sequential(*[f[i] for i in range(7, 11)])
            self.layer3 = nn.Sequential(*[f[i] for i in range(11, 18)])
            self.layer4 = nn.Sequential(f[18])
            stem_ch, low_ch, final_ch, self.hrfp_scale = 16, 32, 1280, 2
        else:
            m = torchvision.models.shufflenet_v2_x1_0(weights=None)
            self.layer0 = _ShuffleLayer0(m.conv1, m.maxpool)
            self.layer1, self.layer2, self.layer3 = m.stage2, m.stage3, m.stage4
            self.layer4 = _ShuffleLayer4(m.conv5)
            stem_ch, low_ch, final_ch, self.hrfp_scale = 24, 116, 1024, 1
        if variant == "D16":                                                # network/deepv3.py:184-187 / :276-279
            for m_ in self.layer3.modules():
                if isinstance(m_, nn.Conv2d) and m_.stride == (2, 2):
                    m_.dilation, m_.padding, m_.stride = (2, 2), (2, 2), (1, 1)
        self.output_stride = 16
        self.aspp = ASPP(final_ch, 256)
        self.bot_fine = nn.Sequential(nn.Conv2d(low_ch, 48, 1, bias=False), nn.BatchNorm2d(48), nn.ReLU(inplace=True))
        self.bot_aspp = nn.Sequential(nn.Conv2d(1280, 256, 1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True))
        self.final1 = nn.Sequential(nn.Conv2d(304, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                                    nn.Conv2d(256, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True))
        self.final2 = nn.Sequential(nn.Conv2d(256, num_classes, 1, bias=True))
        self._build_hrfp(stem_ch)
        init_head(self.aspp, self.bot_aspp, self.bot_fine, self.final1, self.final2)
        self.eps = 1e-5
        self.whitening = False
        self.three_input_layer = False

    def _stem(self, x):
        """layer0 = conv(s) -> ... -> InstanceNorm2d(affine) -> ReLU -> maxpool (Resnet.py:591-599; ResNet3X3: :476-494)."""
        if self.trunk in MOBILE_TRUNKS:
            return self.layer0(x)
        if FUSE_INSTNORM and x.is_cuda and x.dtype == torch.float32:
            mods = list(self.layer0)
            for m in mods[:-3]:
                x = m(x)
            return mods[-1](_instnorm.module_instance_norm_relu(mods[-3], x, True))
        return self.layer0(x)

    def forward(self, x, gts=None, training=True, gates=None):
        """deepv3.py:280-367.  `gates` (extension): the three Bernoulli draws (p, p2, p3) made by the caller with the
        reference's `random.random()` calls instead of here — a CUDA-graph trainer has to know the branch before it
        picks the graph to replay (mrfp_b200/train_step.py)."""
        p, p2, p3 = gates if gates is not None else (random.random(), random.random(), random.random())   # deepv3.py:281-283
        h, w = x.shape[2:]
        he, we = h * self.hrfp_scale, w * self.hrfp_scale                   # the chain is sized relative to xp (stride-2 stems: 2x)
        if training and p < 0.5:
            self.reinit_hrfp()                                              # deepv3.py:290-306
        xp = self._stem(x)                                                  # deepv3.py:309-316
        x, ocout_dec = self.mrfp_stem(xp, he, we, training, p, p2, p3)      # deepv3.py:317-330
        last = self.layer1[-1]
        fused_sums = isinstance(last, Bottleneck)                           # ResNet hosts: layer1's last ReLU leaves the plane sums
        if fused_sums:
            last.emit_plane_sums = bool(training and p2 < 0.5 and self.fuse_layer1_np)
        x = self.layer1(x)                                                  # deepv3.py:332
        if training and p2 < 0.5 and fused_sums and last.np_applied:        # deepv3.py:334-335 already applied inside layer1's last block
            last.np_applied = False
        elif training and p2 < 0.5:                                         # deepv3.py:334-335
            if fused_sums and last.plane_sums is not None:      # statistics came with layer1's last ReLU: NP+ is one streaming pass
                alpha, eps = _npplus.draw_np_plus_factors(x)
                x = _npplus.np_plus_presummed(x, last.plane_sums, alpha, eps)
                last.plane_sums = None
            else:
                x = self.Normalization_Perturbation_Plus(x)
        low_level = x
        x = self.layer4(self.layer3(self.layer2(x)))
        dec0_up = self.bot_aspp(self.aspp(x))
        dec0_fine = self.bot_fine(low_level)
        dec0 = torch.cat([dec0_fine, upsample_bilinear(dec0_up, low_level.shape[2:])], 1)
        dec1 = self.final1(dec0)
        if (training and p3 < 0.5 and self.fuse_plus_tail and self.fuse_final2
                and _hrfp.tail_final2_supported(dec1, self.final2[0], ocout_dec)):
            # deepv3.py:355-361 in one kernel: Upsample + HRFP+ add + the 1x1 classifier (nothing at (N,256,h/2,w/2))
            main_out = upsample_bilinear(_hrfp.hrfp_plus_final2(dec1, self.final2[0], ocout_dec), (h, w))
            return self.criterion(main_out, gts)
        if training and p3 < 0.5:                                           # deepv3.py:355-357
            if self.fuse_plus_tail and isinstance(ocout_dec, _hrfp.HrfpDec) and dec1.is_cuda:
                dec1 = _hrfp.hrfp_plus_add_upsampled(dec1, ocout_dec)       # Upsample + add in one kernel
            else:
                dec1 = upsample_bilinear(dec1, (int(he / 2), int(we / 2)))
                dec1 = self._plus_add(dec1, ocout_dec)
        main_out = upsample_bilinear(self.final2(dec1), (h, w))
        if training:
            return self.criterion(main_out, gts)
        return main_out
