"""PWC-Net as EAVSR's training path uses it -- the caller of the cost volume and of backwarp
(SURVEY.md section 8 rows a6, a7, a8).

Reference: models/pwc_net.py (``PWCNET``: Extractor :29-95, Decoder :97-207 with the correlation call
sites :157-158,165-167 and ``backwarp`` :184-207, Refiner :209-230) and the wrapper in
models/base_model.py (``estimate`` :294-319, ``get_backwarp`` :344-354, ``get_flow`` :356-360).
The module tree is table-driven here but reproduces the reference's parameter names and shapes
(``netExtractor.netOne.0.weight`` ... ``netTwo.netUpfeat.weight`` ... ``netRefiner.netMain.12.bias``), so a
``pwc-default`` state dict (keys ``module*`` -> ``net*`` as at models/pwc_net.py:249-251) loads with
``strict=True``.  The network is frozen and only ever evaluated under ``no_grad`` (models/base_model.py:356-360);
the cost volume and the warps run in libeavsr_b200.so, the dense convolutions stay on cuDNN.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ops import FunctionCorrelation, backwarp, get_backwarp as _get_backwarp

__all__ = ["PWCNET", "estimate", "get_flow", "get_backwarp"]

_LEVEL_NAMES = ("netOne", "netTwo", "netThr", "netFou", "netFiv", "netSix")
_PYRAMID = (3, 16, 32, 64, 96, 128, 196)                 # channels of the input and of the six extractor levels
_DENSE = (128, 128, 96, 64, 32)                           # DenseNet-style decoder widths
# decoder input width per level (index = level): cost volume + features + up-sampled flow + up-sampled feat
_CURRENT = {2: 81 + 32 + 2 + 2, 3: 81 + 64 + 2 + 2, 4: 81 + 96 + 2 + 2, 5: 81 + 128 + 2 + 2, 6: 81}
_BACKWARP_SCALE = {2: 5.0, 3: 2.5, 4: 1.25, 5: 0.625}     # flow magnitude at the level's resolution


def _lrelu():
    return nn.LeakyReLU(negative_slope=0.1, inplace=False)


class _Extractor(nn.Module):
    def __init__(self):
        super().__init__()
        for name, cin, cout in zip(_LEVEL_NAMES, _PYRAMID[:-1], _PYRAMID[1:]):
            setattr(self, name, nn.Sequential(nn.Conv2d(cin, cout, 3, 2, 1), _lrelu(), nn.Conv2d(cout, cout, 3, 1, 1),
                                              _lrelu(), nn.Conv2d(cout, cout, 3, 1, 1), _lrelu()))

    def forward(self, x):
        out = []
        for name in _LEVEL_NAMES:
            x = getattr(self, name)(x)
            out.append(x)
        return out


class _Decoder(nn.Module):
    def __init__(self, level: int):
        super().__init__()
        cur = _CURRENT[level]
        if level < 6:
            self.netUpflow = nn.ConvTranspose2d(2, 2, 4, 2, 1)
            self.netUpfeat = nn.ConvTranspose2d(_CURRENT[level + 1] + sum(_DENSE), 2, 4, 2, 1)
            self.fltBackwarp = _BACKWARP_SCALE[level]
        width = cur
        for name, cout in zip(_LEVEL_NAMES, _DENSE):
            setattr(self, name, nn.Sequential(nn.Conv2d(width, cout, 3, 1, 1), _lrelu()))
            width += cout
        self.netSix = nn.Sequential(nn.Conv2d(width, 2, 3, 1, 1))

    def forward(self, first, second, previous):
        if previous is None:
            feat = F.leaky_relu(FunctionCorrelation(tenFirst=first, tenSecond=second), 0.1)
        else:
            flow = self.netUpflow(previous["tenFlow"])
            up = self.netUpfeat(previous["tenFeat"])
            warped = backwarp(tenInput=second, tenFlow=flow * self.fltBackwarp)
            volume = F.leaky_relu(FunctionCorrelation(tenFirst=first, tenSecond=warped.contiguous()), 0.1)
            feat = torch.cat([volume, first, flow, up], 1)
        for name in _LEVEL_NAMES[:-1]:
            feat = torch.cat([getattr(self, name)(feat), feat], 1)
        return {"tenFlow": self.netSix(feat), "tenFeat": feat}


class _Refiner(nn.Module):
    def __init__(self):
        super().__init__()
        spec = ((_CURRENT[2] + sum(_DENSE), 128, 1), (128, 128, 2), (128, 128, 4), (128, 96, 8), (96, 64, 16),
                (64, 32, 1), (32, 2, 1))
        layers = []
        for i, (cin, cout, d) in enumerate(spec):
            layers.append(nn.Conv2d(cin, cout, 3, 1, d, d))
            if i + 1 < len(spec):
                layers.append(_lrelu())
        self.netMain = nn.Sequential(*layers)

    def forward(self, x):
        return self.netMain(x)


class PWCNET(nn.Module):
    """PWC-Net (models/pwc_net.py:25-261).  Unlike the reference's constructor this one does not read
    ``./pwc/pwc-default`` (the blob is stripped from the reference tree): load a state dict explicitly."""

    def __init__(self):
        super().__init__()
        self.netExtractor = _Extractor()
        for lvl, name in zip((2, 3, 4, 5, 6), _LEVEL_NAMES[1:]):
            setattr(self, name, _Decoder(lvl))
        self.netRefiner = _Refiner()

    def forward(self, tenFirst, tenSecond):
        first, second = self.netExtractor(tenFirst), self.netExtractor(tenSecond)
        est = None
        for k, name in enumerate(reversed(_LEVEL_NAMES[1:]), start=1):      # netSix (coarsest) ... netTwo
            est = getattr(self, name)(first[-k], second[-k], est)
        return est["tenFlow"] + self.netRefiner(est["tenFeat"])


def estimate(tenFirst, tenSecond, net):
    """BaseModel.estimate (models/base_model.py:294-319): resize to multiples of 64, 20 x the network
    output resized back, flow components rescaled."""
    h, w = tenFirst.shape[2:]
    assert tenSecond.shape[2:] == (h, w)
    hp, wp = int(math.ceil(h / 64.0) * 64), int(math.ceil(w / 64.0) * 64)
    a = F.interpolate(tenFirst, size=(hp, wp), mode="bilinear", align_corners=False)
    b = F.interpolate(tenSecond, size=(hp, wp), mode="bilinear", align_corners=False)
    flow = 20.0 * F.interpolate(net(a, b), size=(h, w), mode="bilinear", align_corners=False)
    return flow * flow.new_tensor([w / wp, h / hp]).view(1, 2, 1, 1)


def get_flow(tenFirst, tenSecond, net):
    """BaseModel.get_flow (models/base_model.py:356-360): frozen network, eval mode, no_grad."""
    with torch.no_grad():
        net.eval()
        return estimate(tenFirst, tenSecond, net)


def get_backwarp(tenFirst, tenSecond, net, flow=None, scale=1):
    """BaseModel.get_backwarp (models/base_model.py:344-354): flow from the LR frame to the down-scaled second
    frame, nearest-upsampled x scale, then ``tenSecond`` warped with its validity mask."""
    if flow is None:
        second = F.interpolate(tenSecond, scale_factor=1 / scale, mode="bilinear", align_corners=True)
        flow = get_flow(tenFirst, second, net)
        flow = F.interpolate(flow, scale_factor=scale, mode="nearest") * scale
    return _get_backwarp(tenSecond, flow)
