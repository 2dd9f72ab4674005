#!/usr/bin/env python
"""Device time of the grouped 3x3 convolution kernels (csrc/grouped_conv.cu) at the training shape vs cuDNN."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from eavsr_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


for cin, cout in ((128, 128), (128, 64)):
    conv = torch.nn.Conv2d(cin, cout, 3, 1, 1, groups=cout).to(dev, torch.bfloat16).to(memory_format=torch.channels_last)
    x = torch.randn(8, cin, 64, 64, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_()
    go = torch.randn(8, cout, 64, 64, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    y = ops.grouped_conv3x3(conv, x)
    yt = conv(x)
    f_native = timeit(lambda: ops.grouped_conv3x3(conv, x))
    f_cudnn = timeit(lambda: conv(x))
    b_native = timeit(lambda: torch.autograd.grad(y, [x, conv.weight, conv.bias], go, retain_graph=True))
    b_native_x = timeit(lambda: torch.autograd.grad(y, [x], go, retain_graph=True))
    b_cudnn = timeit(lambda: torch.autograd.grad(yt, [x, conv.weight, conv.bias], go, retain_graph=True))
    print(f"{cin}->{cout} 8x64x64 bf16 (eager, incl. ~25 us host per call): fwd native {f_native:.1f} us, cuDNN {f_cudnn:.1f} us; "
          f"bwd native {b_native:.1f} us (d(input) only {b_native_x:.1f}), cuDNN {b_cudnn:.1f} us")
