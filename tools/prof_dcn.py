#!/usr/bin/env python
"""A few launches of the kernels worth profiling (run under ncu): DCNv2 tcgen05 forward at the bench
shape with sigma=2 and sigma=0.25 offsets, and the 64-channel bf16 warp."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import eavsr_b200 as E  # noqa: E402

dev = torch.device("cuda:0")
dg = int(os.environ.get("DG", "8"))
dt = torch.float32 if os.environ.get("DT", "bf16") == "f32" else torch.bfloat16
h, w = 270, 480
g = torch.Generator().manual_seed(0)
x = torch.randn(1, 64, h, w, generator=g).to(dev, dt).contiguous(memory_format=torch.channels_last)
off = (torch.randn(1, dg * 18, h, w, generator=g) * 2).clamp(-12, 12).to(dev)
msk = torch.sigmoid(torch.randn(1, dg * 9, h, w, generator=g)).to(dev)
wgt = ((torch.rand(64, 64, 3, 3, generator=g) * 2 - 1) / 24).to(dev, dt)
bias = torch.zeros(64, device=dev, dtype=dt)
flow = (torch.randn(1, 2, h, w, generator=g) * 3).to(dev)
for s in (1.0, 0.125):
    o = off * s
    for _ in range(2):
        E.modulated_deform_conv2d(x, o, msk, wgt, bias, 1, 1, 1, 1, dg)
for _ in range(2):
    E.flow_warp(x, flow)
torch.cuda.synchronize()
print("done")
