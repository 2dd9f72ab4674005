#!/usr/bin/env python
"""Timing experiment for the tcgen05 conv: EAVSR_CONV_DBG=1 skips the MMAs (staging + epilogue only),
=2 uses the 128-byte aligned view for every tap (wrong results, tells what mis-aligned A views cost)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from eavsr_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
conv = torch.nn.Conv2d(64, 64, 3, 1, 1).to(dev, torch.bfloat16)
xs = [torch.randn(1, 64, 272, 480, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for _ in range(6)]
for dbg in ("0", "1", "2"):
    os.environ["EAVSR_CONV_DBG"] = dbg
    for sums in (False,):
        with torch.no_grad():
            for i in range(3):
                ops.conv3x3_64(conv, xs[i], 0.0, want_sums=sums)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                keep = [ops.conv3x3_64(conv, xs[i % 6], 0.0, want_sums=sums) for i in range(40)]
            g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            print(f"dbg={dbg} sums={sums}: {a.elapsed_time(b) / 40 * 1e3:.2f} us")
