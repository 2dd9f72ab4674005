#!/usr/bin/env python
"""Per-layer time of the chained tcgen05 convolution (eavsr_conv3x3_chain_forward) against a launch per
convolution: a plain chain of 20 convolutions and a full RCAGroup (30 blocks, 61 convolutions) at 272x480."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import eavsr_b200.model as M  # noqa: E402
from eavsr_b200 import _lib as L  # noqa: E402
from eavsr_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
h, w = int(os.environ.get("H", 272)), int(os.environ.get("W", 480))
n = int(os.environ.get("N", 1))


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


with torch.no_grad():
    x = torch.randn(n, 64, h, w, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    convs = [torch.nn.Conv2d(64, 64, 3, 1, 1).to(dev, torch.bfloat16) for _ in range(20)]
    bufs = [torch.empty_like(x), torch.empty_like(x)]
    lib = L.load()
    sync = torch.zeros(1, dtype=torch.int32, device=dev)
    biases = [c.bias.detach().contiguous() for c in convs]

    def plain_chain(k):
        arr = (L.ConvLayer * k)()
        src = x
        for i in range(k):
            a = arr[i]
            a.x, a.out = src.data_ptr(), bufs[i & 1].data_ptr()
            a.packed_weight = ops._packed_conv_weight(convs[i], dev).data_ptr()
            a.bias = biases[i].data_ptr()
            a.negative_slope = 0.0
            src = bufs[i & 1]
        L.check(lib.eavsr_conv3x3_chain_forward(arr, k, n, h, w, L.BF16, sync.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream), "chain")

    def per_launch(k):
        src = x
        for i in range(k):
            src = ops.conv3x3_64(convs[i], src, 0.0)

    res = {}
    t1, t20 = timeit(lambda: plain_chain(1)), timeit(lambda: plain_chain(20))
    res["plain conv, chain of 1 (us)"] = round(t1, 2)
    res["plain conv, chain of 20 (us per layer)"] = round(t20 / 20, 2)
    res["plain conv, 20 launches (us per layer)"] = round(timeit(lambda: per_launch(20)) / 20, 2)
    grp = M._RCAGroup(64, 30).to(dev, torch.bfloat16).to(memory_format=torch.channels_last)
    grp.chain = True
    res["RCAGroup(30) chained (us per conv)"] = round(timeit(lambda: grp(x), 5) / 61, 2)
    grp.chain = False
    res["RCAGroup(30) launch per conv (us per conv)"] = round(timeit(lambda: grp(x), 5) / 61, 2)
    res["shape"] = [n, 64, h, w]
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/prof_chain.json", "w"), indent=1)

# per-phase timestamps of CTA 0 (only with a library built with -DEAVSR_CONV_TRACE)
import ctypes  # noqa: E402
lib = L.load()
if hasattr(lib, "eavsr_debug_conv_trace"):
    with torch.no_grad():
        plain_chain(6)
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * (6 * 16))()
        lib.eavsr_debug_conv_trace(buf, 6 * 16)
        names = ["layer start", "after grid barrier (C)", "after setup (D)", "weights landed", "tile0 staged", "tile1 staged",
                 "last tile staged", "last MMA issued", "epi: first acc ready", "epi: last acc ready", "epi: done"]
        t0 = buf[0]
        for li in range(6):
            row = [buf[li * 16 + k] for k in range(11)]
            print(f"layer {li}: " + "  ".join(f"{nm}={(v - t0) / 1965.0:.2f}us" for nm, v in zip(names, row)))
        grp.chain = True
        grp(x)
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * (8 * 16))()
        lib.eavsr_debug_conv_trace(buf, 8 * 16)
        t0 = buf[0]
        for li in range(8):
            row = [buf[li * 16 + k] for k in range(11)]
            print(f"rca layer {li}: " + "  ".join(f"{nm}={(v - t0) / 1965.0:.2f}us" for nm, v in zip(names, row)))

        if hasattr(lib, "eavsr_debug_conv_trace2"):
            b2 = (ctypes.c_ulonglong * 512)()
            lib.eavsr_debug_conv_trace2(b2, 512)
            for li in (2, 3):
                print(f"rca layer {li} per tile (us since layer's first stamp): landed | staged | MMAs issued | acc ready | acc drained | stored")
                base = min(v for v in b2[li * 64:(li + 1) * 64] if v)
                for tl in range(8):
                    row = b2[(li * 8 + tl) * 8:(li * 8 + tl) * 8 + 6]
                    print(f"  tile {tl}: " + "  ".join(f"{(v - base) / 1965.0:6.2f}" if v else "   -  " for v in row))
