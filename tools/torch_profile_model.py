#!/usr/bin/env python
"""torch.profiler view of one eager EAVSR+ x4 forward (which aten ops are left between our kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402

T = int(os.environ.get("FRAMES", "6"))
bench.T_FRAMES = T
wl = bench.ModelWorkload(torch.device("cuda:0"), t=T, graph=False, clips_per_step=int(os.environ.get("CLIPS", "1")))
with torch.no_grad():
    for _ in range(2):
        wl.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        wl.step()
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=40, max_name_column_width=70))
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=60, max_name_column_width=40))
