#!/usr/bin/env python
"""A few launches of the tcgen05 conv3x3 and the fused glue kernels (run under ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from eavsr_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
conv = torch.nn.Conv2d(64, 64, 3, 1, 1).to(dev, torch.bfloat16)
x = torch.randn(1, 64, 272, 480, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
with torch.no_grad():
    for _ in range(3):
        ops.conv3x3_64(conv, x, 0.0)
    for _ in range(2):
        ops.conv3x3_64(conv, x, 1.0, want_sums=True)
torch.cuda.synchronize()
print("done")
