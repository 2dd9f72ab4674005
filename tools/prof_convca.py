#!/usr/bin/env python
"""A few launches of the fused-input conv3x3 (run under ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import eavsr_b200.model as M  # noqa: E402
from eavsr_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
conv = torch.nn.Conv2d(64, 64, 3, 1, 1).to(dev, torch.bfloat16)
du = M._CALayer(64).to(dev, torch.bfloat16).conv_du
mk = lambda: torch.randn(1, 64, 272, 480, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)  # noqa: E731
skip, res = mk(), mk()
sums = torch.randn(1, 64, device=dev) * 100
with torch.no_grad():
    for _ in range(4):
        ops.conv3x3_64_ca(conv, skip, res, sums, du[0].weight, du[0].bias, du[2].weight, du[2].bias, 0.0)
torch.cuda.synchronize()
