#!/usr/bin/env python
"""torch.profiler view of one eager EAVSR+ x4 training step (BASELINE config 5 shapes, bf16 autocast, 1 GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402

wl = bench.TrainWorkload(torch.device("cuda:0"), graph=False, dtype=torch.bfloat16)
for _ in range(3):
    wl.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    wl.step()
    torch.cuda.synchronize()
if os.environ.get("BY_SHAPE"):
    print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=45))
else:
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=90))
