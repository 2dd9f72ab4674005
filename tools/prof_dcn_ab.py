#!/usr/bin/env python
"""A/B timing of the DCN forward generations at the bench shape (graph replay, CUDA events)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from eavsr_b200 import _lib as L  # noqa: E402
from eavsr_b200.ops import _ModulatedDeformConv2dFn  # noqa: E402

dev = torch.device("cuda:0")
h, w, dg = 270, 480, 8
g = torch.Generator().manual_seed(0)
nb = 4
xs = [torch.randn(1, 64, h, w, generator=g).to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last) for _ in range(nb)]
res = {}
for sigma in (2.0, 0.5):
    offs = [(torch.randn(1, dg * 18, h, w, generator=g) * sigma).clamp(-12, 12).to(dev) for _ in range(nb)]
    msks = [torch.sigmoid(torch.randn(1, dg * 9, h, w, generator=g)).to(dev) for _ in range(nb)]
    wgt = ((torch.rand(64, 64, 3, 3, generator=g) * 2 - 1) / 24).to(dev, torch.bfloat16)
    bias = torch.zeros(64, device=dev, dtype=torch.bfloat16)
    for name, flags in (("win3", 0), ("win2", L.DCN_FORCE_WIN2), ("win1", L.DCN_FORCE_WIN1)):
        with torch.no_grad():
            call = lambda i: _ModulatedDeformConv2dFn.apply(xs[i % nb], offs[i % nb], msks[i % nb], wgt, bias, 1, 1, 1, 1, dg, flags)  # noqa: E731
            for i in range(3):
                call(i)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                keep = [call(i) for i in range(40)]
            gr.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            gr.replay()
            b.record()
            torch.cuda.synchronize()
            res[f"{name} sigma={sigma}"] = round(a.elapsed_time(b) * 1e3 / 40 - 0.0, 2)   # incl. the 2 us weight pack per call
            del keep, gr
print(json.dumps(res, indent=1))
