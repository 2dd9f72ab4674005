#!/usr/bin/env python
"""Minimal launch targets for `ncu --set full` captures (one kernel family per invocation).
usage: python tools/ncu_target.py corr|dcn|warp|dcn_bwd"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import eavsr_b200 as E  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "corr"
g = torch.Generator().manual_seed(0)
if what == "corr":
    a = torch.randn(30, 32, 80, 128, generator=g).to(dev)
    b = torch.randn(30, 32, 80, 128, generator=g).to(dev)
    for _ in range(3):
        E.FunctionCorrelation(tenFirst=a, tenSecond=b)
elif what in ("dcn", "dcn_bwd", "dcn_bwd_nox"):
    h, w, dg = 270, 480, 8
    x = torch.randn(1, 64, h, w, generator=g).to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    off = (torch.randn(1, dg * 18, h, w, generator=g) * 2).clamp(-12, 12).to(dev)
    msk = torch.sigmoid(torch.randn(1, dg * 9, h, w, generator=g)).to(dev)
    wgt = ((torch.rand(64, 64, 3, 3, generator=g) * 2 - 1) / 24).to(dev, torch.bfloat16)
    bias = torch.zeros(64, device=dev, dtype=torch.bfloat16)
    if what == "dcn":
        for _ in range(3):
            E.modulated_deform_conv2d(x, off, msk, wgt, bias, 1, 1, 1, 1, dg)
    else:
        x.requires_grad_(what == "dcn_bwd"); off.requires_grad_(); msk.requires_grad_(); wgt.requires_grad_(); bias.requires_grad_()
        for _ in range(2):
            out = E.modulated_deform_conv2d(x, off, msk, wgt, bias, 1, 1, 1, 1, dg)
            out.backward(torch.ones_like(out))
elif what == "conv":
    from eavsr_b200 import ops
    h, w = (272, 480) if len(sys.argv) < 3 else (int(sys.argv[2]), int(sys.argv[3]))
    conv = torch.nn.Conv2d(64, 64, 3, padding=1).to(dev, torch.bfloat16).to(memory_format=torch.channels_last)
    xs = [torch.randn(1, 64, h, w, generator=g).to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last) for _ in range(3)]
    with torch.no_grad():
        for i in range(3):
            ops.conv3x3_64(conv, xs[i], 0.0)
elif what == "warp_bwd":
    x = torch.randn(1, 64, 270, 480, generator=g).to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_()
    flow = (torch.randn(1, 2, 270, 480, generator=g) * 3).to(dev).requires_grad_()
    for _ in range(3):
        out = E.flow_warp(x, flow)
        out.backward(torch.ones_like(out))
    a = torch.randn(8, 32, 80, 128, generator=g).to(dev).requires_grad_()
    b = torch.randn(8, 32, 80, 128, generator=g).to(dev).requires_grad_()
    for _ in range(2):
        o = E.FunctionCorrelation(tenFirst=a, tenSecond=b)
        o.backward(torch.ones_like(o))
elif what == "warp":
    x = torch.randn(8, 64, 270, 480, generator=g).to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    flow = (torch.randn(8, 2, 270, 480, generator=g) * 3).to(dev)
    for _ in range(3):
        E.flow_warp(x, flow)
torch.cuda.synchronize()
