#!/usr/bin/env python
"""torch.cat(dim=1) of channels_last tensors vs eavsr_b200.ops.cat_channels (eager, CUDA events)."""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eavsr_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters
with torch.no_grad():
    for n in (1, 2, 8):
        for k in (2, 3, 5):
            ts = [torch.randn(n, 64, 272, 480, device=dev).bfloat16().contiguous(memory_format=torch.channels_last) for _ in range(k)]
            mb = 2 * k * n * 64 * 272 * 480 * 2 / 1e6
            t0 = timeit(lambda: torch.cat(ts, 1)); t1 = timeit(lambda: ops.cat_channels(ts))
            print(f"n={n} k={k}: torch.cat {t0:.1f} us ({mb / t0 * 1e3:.0f} GB/s)  nhwc_cat {t1:.1f} us ({mb / t1 * 1e3:.0f} GB/s)")
