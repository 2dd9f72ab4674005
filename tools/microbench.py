#!/usr/bin/env python
"""Per-kernel micro-benchmarks on one B200 (SURVEY.md 8d shapes): achieved GB/s / TFLOP/s vs the
measured peaks.  Inputs rotate over > L2-sized pools; timing with CUDA events after warm-up.
Writes gpurun_out/microbench.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import eavsr_b200 as E  # noqa: E402
from bench import load_peaks  # noqa: E402

dev = torch.device("cuda:0")
peaks = load_peaks()
res = []
# Everything -- input creation, forwards kept for the backward timings, graph capture, events -- runs on ONE side
# stream: autograd runs a backward node on the stream its forward ran on, and a forward left on the legacy default
# stream cannot be joined from a capturing stream (cudaErrorStreamCaptureImplicit).
SIDE = torch.cuda.Stream(dev)
torch.cuda.set_stream(SIDE)


def timeit(fn, iters, warm=3, graph=True):
    """Device time per call.  The Python/ctypes wrapper costs ~25 us per call on the host, more than
    most of these kernels take, so the calls are captured into one CUDA graph and the replay is
    timed (CUDA events, after a warm-up replay)."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph:
        g = torch.cuda.CUDAGraph()
        keep = []
        with torch.cuda.graph(g, stream=SIDE):
            for i in range(iters):
                keep.append(fn(i))
        g.replay()
        torch.cuda.synchronize()
        a.record()
        g.replay()
        b.record()
    else:
        a.record()
        for i in range(iters):
            fn(i)
        b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 1e3 / iters


def pool(make, total_bytes, each_bytes, lo=2, hi=12):
    n = max(lo, min(hi, int(total_bytes // max(1, each_bytes)) + 1))
    return [make() for _ in range(n)]


def rec(name, sec, bytes_alg, flops=0.0, **kw):
    r = {"name": name, "us": round(sec * 1e6, 2), "GBps": round(bytes_alg / sec / 1e9, 1),
         "hbm_frac": round(bytes_alg / sec / 1e9 / peaks["hbm_gbs"], 4)}
    if flops:
        r["TFLOPs"] = round(flops / sec / 1e12, 2)
        r["tc_frac"] = round(flops / sec / 1e12 / peaks["bf16_tflops"], 4)
    r.update(kw)
    res.append(r)
    print(json.dumps(r), flush=True)


def bench_warp():
    for dt, es in ((torch.bfloat16, 2), (torch.float32, 4)):
        for n, h, w in ((1, 270, 480), (8, 270, 480), (1, 135, 240), (1, 67, 120)):
            each = n * 64 * h * w * es
            xs = pool(lambda: torch.randn(n, 64, h, w, device=dev).to(dt).contiguous(memory_format=torch.channels_last),
                      400e6, each)
            fl = pool(lambda: torch.randn(n, 2, h, w, device=dev) * 3, 0, 1, lo=len(xs), hi=len(xs))
            sec = timeit(lambda i: E.flow_warp(xs[i % len(xs)], fl[i % len(xs)]), 100)
            rec(f"flow_warp fwd {dt} {n}x64x{h}x{w}", sec, n * h * w * (2 * 64 * es + 8))
            m = min(4, len(xs))
            xg = [x.clone().requires_grad_() for x in xs[:m]]
            fg = [f.clone().requires_grad_() for f in fl[:m]]
            outs = [E.flow_warp(a, b) for a, b in zip(xg, fg)]
            gs = [torch.randn_like(o) for o in outs]
            # device time of the whole backward (memset + scatter kernel + cast), captured into a CUDA graph: the
            # eager number of round 1 was ~100 us of host-side autograd glue per call, not the kernels
            sec = timeit(lambda i: torch.autograd.grad(outs[i % m], [xg[i % m], fg[i % m]], gs[i % m], retain_graph=True), 30)
            rec(f"flow_warp bwd(x,flow) {dt} {n}x64x{h}x{w}", sec, n * h * w * (3 * 64 * es + 16))
    x2 = torch.randn(1, 2, 270, 480, device=dev)
    f2 = torch.randn(1, 270, 480, 2, device=dev)
    sec = timeit(lambda i: E.flow_warp_nhw2(x2, f2), 100)
    rec("flow_warp fwd f32 1x2x270x480 (flow composition, NCHW)", sec, 270 * 480 * (2 * 2 * 4 + 8))
    x3 = torch.rand(29, 3, 288, 480, device=dev)
    f3 = torch.randn(29, 288, 480, 2, device=dev)
    sec = timeit(lambda i: E.flow_warp_nhw2(x3, f3, padding_mode="border"), 50)
    rec("flow_warp fwd f32 29x3x288x480 border (SPyNet, NCHW)", sec, 29 * 288 * 480 * (2 * 3 * 4 + 8))


def bench_dcn():
    from eavsr_b200.ops import _ModulatedDeformConv2dFn
    from eavsr_b200 import _lib as L
    h, w = 270, 480
    for dt, es in ((torch.bfloat16, 2), (torch.float32, 4)):
        for dg in (8, 16):
            for n in (1, 4):
                each = n * h * w * (dg * 27 * 4)
                k = max(2, min(6, int(500e6 // each) + 1))
                xs = [torch.randn(n, 64, h, w, device=dev).to(dt).contiguous(memory_format=torch.channels_last) for _ in range(k)]
                offs = [(torch.randn(n, dg * 18, h, w, device=dev) * 2).clamp(-12, 12) for _ in range(k)]
                msks = [torch.sigmoid(torch.randn(n, dg * 9, h, w, device=dev)) for _ in range(k)]
                wgt = ((torch.rand(64, 64, 3, 3, device=dev) * 2 - 1) / 24).to(dt)
                bias = torch.zeros(64, device=dev, dtype=dt)
                px = n * h * w
                by = px * (128 * es + dg * 27 * 4)
                fl = 2.0 * px * 64 * 64 * 9
                sec = timeit(lambda i: E.modulated_deform_conv2d(xs[i % k], offs[i % k], msks[i % k], wgt, bias, 1, 1, 1, 1, dg), 40)
                rec(f"dcn fwd tc {dt} dg={dg} {n}x64x{h}x{w} sigma=2", sec, by, fl)
                if n == 1:
                    small = [o * 0.25 for o in offs]
                    sec = timeit(lambda i: E.modulated_deform_conv2d(xs[i % k], small[i % k], msks[i % k], wgt, bias, 1, 1, 1, 1, dg), 40)
                    rec(f"dcn fwd tc {dt} dg={dg} {n}x64x{h}x{w} sigma=0.5", sec, by, fl)
                    del small
                if n == 1:
                    sec = timeit(lambda i: _ModulatedDeformConv2dFn.apply(xs[i % k], offs[i % k], msks[i % k], wgt, bias, 1, 1, 1, 1, dg, L.DCN_FORCE_GENERIC), 3, warm=1)
                    rec(f"dcn fwd generic {dt} dg={dg} {n}x64x{h}x{w}", sec, by, fl)
                    for flags, tag in ((0, "tc (data+weight)" if dt == torch.bfloat16 else "generic"),
                                       (L.DCN_BWD_GENERIC_DATA, "generic data + tc weight"),
                                       (L.DCN_BWD_GENERIC_WEIGHT, "tc data + generic weight"),
                                       (L.DCN_BWD_GENERIC_DATA | L.DCN_BWD_GENERIC_WEIGHT, "generic")):
                        if dt != torch.bfloat16 and flags:
                            continue
                        xg = xs[0].clone().requires_grad_()
                        og, mg, wg, bg = offs[0].clone().requires_grad_(), msks[0].clone().requires_grad_(), wgt.clone().requires_grad_(), bias.clone().requires_grad_()
                        out = _ModulatedDeformConv2dFn.apply(xg, og, mg, wg, bg, 1, 1, 1, 1, dg, flags)
                        go = torch.randn_like(out)
                        sec = timeit(lambda i: torch.autograd.grad(out, [xg, og, mg, wg, bg], go, retain_graph=True), 5, warm=2)
                        rec(f"dcn bwd {tag} {dt} dg={dg} {n}x64x{h}x{w} (graph replay)", sec, by * 2, fl * 2)
                del xs, offs, msks
                torch.cuda.empty_cache()


def bench_dcn_affine():
    """row f1: DCN with the offset expansion fused in vs expansion kernel + DCN."""
    from eavsr_b200 import ops
    h, w, D = 270, 480, 8
    k = 4
    xs = [torch.randn(1, 64, h, w, device=dev).bfloat16().contiguous(memory_format=torch.channels_last) for _ in range(k)]
    affs = []
    for _ in range(k):
        a = torch.randn(1, 15 * D, h, w, device=dev)
        a[:, :4 * D] = a[:, :4 * D] * 0.2 + torch.tensor([1.0, 0.0, 0.0, 1.0], device=dev).repeat(D).view(1, -1, 1, 1)
        affs.append(a.bfloat16().contiguous(memory_format=torch.channels_last))
    ab = (torch.randn(15 * D, device=dev) * 0.1).bfloat16()
    wgt = ((torch.rand(64, 64, 3, 3, device=dev) * 2 - 1) / 24).bfloat16()
    bias = torch.zeros(64, device=dev, dtype=torch.bfloat16)
    px = h * w
    fl = 2.0 * px * 64 * 64 * 9
    with torch.no_grad():
        sec = timeit(lambda i: ops.dcn_affine(xs[i % k], affs[i % k], ab, wgt, bias, D, static_weight=True), 40)
        rec(f"dcn_affine (fused offsets) bf16 dg=8 1x64x{h}x{w}", sec, px * (128 + 240 + 128), fl)

        def two(i):
            a = affs[i % k]
            off, msk = ops.affine_offsets_mask(a[:, :4 * D], a[:, 4 * D:6 * D], a[:, 6 * D:], D, ab[:4 * D], ab[4 * D:6 * D], ab[6 * D:])
            return E.modulated_deform_conv2d(xs[i % k], off, msk, wgt, bias, 1, 1, 1, 1, D, static_weight=True)
        sec = timeit(two, 40)
        rec(f"affine_offsets + dcn (unfused) bf16 dg=8 1x64x{h}x{w}", sec, px * (128 + 240 + 128), fl)


def bench_corr():
    fwd_only = os.environ.get("CORR_FWD_ONLY") == "1"
    for n, c, h, w in ((30, 32, 80, 128), (30, 64, 40, 64), (30, 96, 20, 32), (30, 128, 10, 16), (30, 196, 5, 8),
                       (8, 32, 16, 16), (8, 196, 1, 1)):
        each = 2 * n * c * h * w * 4
        k = max(2, min(8, int(300e6 // each) + 1))
        a = [torch.randn(n, c, h, w, device=dev) for _ in range(k)]
        b = [torch.randn(n, c, h, w, device=dev) for _ in range(k)]
        sec = timeit(lambda i: E.FunctionCorrelation(tenFirst=a[i % k], tenSecond=b[i % k]), 50)
        rec(f"correlation fwd f32 {n}x{c}x{h}x{w}", sec, n * h * w * (2 * c + 81) * 4, 2.0 * 81 * c * n * h * w)
        if fwd_only:
            continue
        ag, bg = a[0].clone().requires_grad_(), b[0].clone().requires_grad_()
        out = E.FunctionCorrelation(tenFirst=ag, tenSecond=bg)
        go = torch.randn_like(out)
        sec = timeit(lambda i: torch.autograd.grad(out, [ag, bg], go, retain_graph=True), 10, graph=False)
        rec(f"correlation bwd f32 {n}x{c}x{h}x{w}", sec, n * h * w * (4 * c + 81) * 4, 4.0 * 81 * c * n * h * w)


def bench_conv():
    from eavsr_b200 import ops
    import torch.nn.functional as F
    for n, h, w in ((1, 272, 480), (1, 1088, 1920)):
        k = 6 if h < 1000 else 2
        conv = torch.nn.Conv2d(64, 64, 3, 1, 1).to(dev, torch.bfloat16).to(memory_format=torch.channels_last)
        xs = [torch.randn(n, 64, h, w, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for _ in range(k)]
        px = n * h * w
        by, fl = px * 256, 2.0 * px * 64 * 64 * 9
        with torch.no_grad():
            sec = timeit(lambda i: ops.conv3x3_64(conv, xs[i % k], 0.0), 40)
            rec(f"conv3x3 64->64 tcgen05 +bias+relu bf16 {n}x64x{h}x{w}", sec, by, fl)
            sec = timeit(lambda i: ops.conv3x3_64(conv, xs[i % k], 1.0, want_sums=True), 40)
            rec(f"conv3x3 64->64 tcgen05 +bias+chan-sums bf16 {n}x64x{h}x{w}", sec, by, fl)
            sec = timeit(lambda i: F.relu(conv(xs[i % k])), 40)
            rec(f"cuDNN conv3x3 + bias + relu (torch) bf16 {n}x64x{h}x{w}", sec, by, fl)
            sec = timeit(lambda i: F.conv2d(xs[i % k], conv.weight, None, 1, 1), 40)
            rec(f"cuDNN conv3x3 only (torch) bf16 {n}x64x{h}x{w}", sec, by, fl)


def bench_convca():
    """conv3x3 with the previous block's channel attention fused into its loader vs ca_scale + conv3x3."""
    from eavsr_b200 import ops
    import eavsr_b200.model as M
    h, w = 272, 480
    conv = torch.nn.Conv2d(64, 64, 3, 1, 1).to(dev, torch.bfloat16)
    du = M._CALayer(64).to(dev, torch.bfloat16).conv_du
    k = 6
    mk = lambda: torch.randn(1, 64, h, w, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)   # noqa: E731
    skips, ress = [mk() for _ in range(k)], [mk() for _ in range(k)]
    sums = torch.randn(1, 64, device=dev) * 100
    by = h * w * 64 * 2 * 4
    fl = 2.0 * h * w * 64 * 64 * 9
    with torch.no_grad():
        sec = timeit(lambda i: ops.conv3x3_64_ca(conv, skips[i % k], ress[i % k], sums, du[0].weight, du[0].bias, du[2].weight, du[2].bias, 0.0), 40)
        rec(f"conv3x3 + fused channel attention input bf16 1x64x{h}x{w}", sec, by, fl)
        sec = timeit(lambda i: ops.conv3x3_64(conv, ops.ca_scale(ress[i % k], skips[i % k], sums, du[0].weight, du[0].bias, du[2].weight, du[2].bias, 16), 0.0), 40)
        rec(f"ca_scale + conv3x3 (2 launches) bf16 1x64x{h}x{w}", sec, by, fl)


if __name__ == "__main__":
    which = sys.argv[1:] or ["warp", "dcn", "corr", "conv"]
    for wname in which:
        globals()["bench_" + wname]()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"peaks": peaks, "results": res}, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)
