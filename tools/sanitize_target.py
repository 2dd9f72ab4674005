"""Unit shapes of every hot-path operator, for `compute-sanitizer --tool memcheck|racecheck|initcheck`
(SURVEY.md section 4).  Run through tools/sanitize.sh on the GPU box; the logs go to profiles/."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch  # noqa: E402

import eavsr_b200 as E  # noqa: E402
from eavsr_b200 import ops  # noqa: E402
from helpers import dcn_inputs, warp_inputs  # noqa: E402

dev = torch.device("cuda:0")
cl = lambda t: t.contiguous(memory_format=torch.channels_last)      # noqa: E731
which = sys.argv[1:] or ["warp", "dcn", "corr", "fused"]

if "warp" in which:
    for dtype in (torch.float32, torch.bfloat16):
        for shape in ((1, 64, 19, 37), (2, 3, 9, 15), (1, 20, 17, 19)):
            x, flow = warp_inputs(*shape, seed=1)
            xg = x.to(dev, dtype)
            xg = (cl(xg) if shape[1] % 4 == 0 else xg).requires_grad_()
            fg = flow.to(dev).requires_grad_()
            for pad in ("zeros", "border"):
                E.flow_warp(xg, fg, padding_mode=pad).float().sum().backward()
    x, flow = warp_inputs(1, 3, 12, 16, seed=2)
    E.get_backwarp(x.to(dev).requires_grad_(), flow.to(dev))[0].sum().backward()
    a, b = (cl(torch.randn(1, 64, 21, 33, device=dev).bfloat16()) for _ in range(2))
    with torch.no_grad():
        ops.flow_warp2(a, b, torch.randn(1, 2, 21, 33, device=dev))
        c, d = (cl(torch.randn(1, 64, 20, 32, device=dev).bfloat16()) for _ in range(2))
        ops.flow_warp_pyramid(c, [(torch.randn(1, 2, 80, 128, device=dev), 0.25)])
        ops.flow_warp_pyramid(c, [(torch.randn(1, 2, 20, 32, device=dev), 1.0), (torch.randn(1, 2, 10, 16, device=dev), 2.0)],
                              x2=d, want_flow=True, keep=(1,))
        ops.spynet_level_input(torch.rand(2, 3, 18, 30, device=dev), torch.rand(2, 3, 18, 30, device=dev),
                               torch.randn(2, 2, 9, 15, device=dev))
if "dcn" in which:
    for dg, dtype in ((8, torch.bfloat16), (8, torch.float32), (16, torch.bfloat16), (4, torch.bfloat16)):
        x, off, mask, w, b = dcn_inputs(1, 64, 21, 35, 64, dg, seed=3)
        leaves = [cl(x.to(dev, dtype)).requires_grad_(), off.to(dev).requires_grad_(), mask.to(dev).requires_grad_(),
                  w.to(dev, dtype).requires_grad_(), b.to(dev, dtype).requires_grad_()]
        E.modulated_deform_conv2d(*leaves, 1, 1, 1, 1, dg).float().square().mean().backward()
    # w % 4 == 0, bf16, dg = 8: the TMA-fed window kernels (fifth generation, and the fourth via its force flag)
    from eavsr_b200 import _lib as L
    from eavsr_b200.ops import _ModulatedDeformConv2dFn
    x, off, mask, w, b = dcn_inputs(2, 64, 21, 36, 64, 8, seed=5)
    args = (cl(x.to(dev, torch.bfloat16)), off.to(dev), mask.to(dev), w.to(dev, torch.bfloat16), b.to(dev, torch.bfloat16))
    with torch.no_grad():
        for fl in (0, L.DCN_FORCE_WIN2, L.DCN_BLEND_FP32):
            _ModulatedDeformConv2dFn.apply(*args, 1, 1, 1, 1, 8, fl)
    x, off, mask, w, b = dcn_inputs(1, 12, 9, 11, 8, 3, seed=4, groups=2)
    leaves = [t.to(dev).requires_grad_() for t in (x, off, mask, w, b)]
    E.modulated_deform_conv2d(*leaves, 1, 1, 1, 2, 3).square().mean().backward()
if "corr" in which:
    for shape in ((1, 32, 24, 32), (2, 7, 9, 13), (2, 64, 4, 4)):
        f1 = torch.randn(shape, device=dev, requires_grad=True)
        f2 = torch.randn(shape, device=dev, requires_grad=True)
        E.FunctionCorrelation(tenFirst=f1, tenSecond=f2).square().mean().backward()
    ops.CORRELATION_TF32 = True                  # the opt-in tcgen05 / tf32 banded GEMM
    for shape in ((1, 32, 24, 32), (2, 20, 24, 72), (1, 5, 18, 16)):
        E.FunctionCorrelation(tenFirst=torch.randn(shape, device=dev), tenSecond=torch.randn(shape, device=dev))
    ops.CORRELATION_TF32 = False
if "fused" in which:
    from eavsr_b200.model import MultiAdSTN, _RCAGroup
    from eavsr_b200.synthetic import seeded_parameters
    with torch.no_grad():
        m = MultiAdSTN(64, 8)
        seeded_parameters(m)
        m = m.to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        feats = lambda: [cl(torch.randn(1, 64, 24 >> i, 40 >> i, device=dev).bfloat16()) for i in range(3)]   # noqa: E731
        m(feats(), feats(), feats()[0], torch.randn(1, 2, 24, 40, device=dev))
        g = _RCAGroup(64, 2)
        seeded_parameters(g)
        g = g.to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        g(feats()[0])                            # chained cooperative launch
        g.chain = False
        g(feats()[0])                            # one launch per convolution
        xs = [cl(torch.randn(2, c, 9, 13, device=dev).bfloat16()) for c in (64, 8, 128)]
        ops.cat_channels(xs)                     # nhwc_cat, whole buffer and channel slice
        buf = torch.zeros(2, 256, 9, 13, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ops.cat_channels(xs, out=buf, channel_offset=16)
torch.cuda.synchronize()
print("sanitize_target done:", which)
