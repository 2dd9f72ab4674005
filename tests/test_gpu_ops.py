"""GPU parity tests: the CUDA path (through the ctypes/C-ABI binding) against the CPU oracle on
the same seeded inputs.  Tolerances: fp32 path max-abs <= 1e-3 (BASELINE.json north_star; the
kernels are far inside it), bf16 path relative to the output RMS."""
import os

import numpy as np
import pytest
import torch

import eavsr_b200 as E
from eavsr_b200 import _lib as L
from eavsr_b200.ops import _ModulatedDeformConv2dFn
from oracle import alignment as O

from helpers import dcn_inputs, max_err, rel_err, smooth_flow_mask, warp_inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

FP32_TOL = 1e-3     # north_star: max abs error <= 1e-3 fp32
BF16_REL = 3e-2     # bf16 storage rounding (2^-9) relative to output RMS, worst element


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


# ---------------------------------------------------------------------------------------------
# flow_warp
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", ["n2hw", "nhw2"])
@pytest.mark.parametrize("pad", ["zeros", "border"])
@pytest.mark.parametrize("shape,cl", [((2, 64, 33, 47), True), ((1, 64, 67, 120), True), ((2, 64, 16, 24), False),
                                      ((3, 2, 31, 45), False), ((2, 3, 9, 15), False), ((1, 20, 17, 19), True),
                                      ((1, 196, 5, 8), False), ((1, 32, 1, 1), True)])
def test_flow_warp_fp32(cuda, layout, pad, shape, cl):
    x, flow = warp_inputs(*shape, seed=1, layout=layout)
    ref = O.flow_warp(x.double(), flow.double(), layout, pad)
    xg = x.to(cuda)
    xg = _cl(xg) if cl else xg
    fn = E.flow_warp if layout == "n2hw" else E.flow_warp_nhw2
    out = fn(xg, flow.to(cuda), padding_mode=pad)
    assert out.shape == ref.shape
    assert max_err(out, ref) < 1e-4


@pytest.mark.parametrize("pad", ["zeros", "border"])
def test_flow_warp_bf16(cuda, pad):
    x, flow = warp_inputs(2, 64, 40, 56, seed=2)
    xb = x.bfloat16()
    ref = O.flow_warp(xb.double(), flow.double(), "n2hw", pad)
    out = E.flow_warp(_cl(xb.to(cuda)), flow.to(cuda), padding_mode=pad)
    assert out.dtype == torch.bfloat16
    assert rel_err(out, ref) < BF16_REL


def test_flow_warp_identity_and_integer_shift(cuda):
    x = torch.randn(1, 64, 20, 30)
    flow = torch.zeros(1, 2, 20, 30)
    out = E.flow_warp(_cl(x.to(cuda)), flow.to(cuda))
    assert torch.equal(out.cpu(), x)
    flow[:, 0] = 3.0
    flow[:, 1] = -2.0
    out = E.flow_warp(_cl(x.to(cuda)), flow.to(cuda)).cpu()
    exp = torch.zeros_like(x)
    exp[:, :, 2:, :27] = x[:, :, :18, 3:]
    assert torch.equal(out, exp)


@pytest.mark.parametrize("layout", ["n2hw", "nhw2"])
@pytest.mark.parametrize("pad", ["zeros", "border"])
@pytest.mark.parametrize("shape,cl", [((2, 64, 19, 23), True), ((2, 2, 21, 17), False), ((1, 24, 9, 11), True),
                                      ((1, 12, 7, 9), True)])
def test_flow_warp_backward(cuda, layout, pad, shape, cl):
    x, flow = warp_inputs(*shape, seed=3, layout=layout, sigma=2.0)
    g = torch.randn(shape, generator=torch.Generator().manual_seed(4))
    xr = x.double().requires_grad_()
    fr = flow.double().requires_grad_()
    ref = O.flow_warp(xr, fr, layout, pad)
    gx_ref, gf_ref = torch.autograd.grad(ref, [xr, fr], g.double())
    xg = (_cl(x.to(cuda)) if cl else x.to(cuda)).requires_grad_()
    fg = flow.to(cuda).requires_grad_()
    fn = E.flow_warp if layout == "n2hw" else E.flow_warp_nhw2
    out = fn(xg, fg, padding_mode=pad)
    gx, gf = torch.autograd.grad(out, [xg, fg], g.to(cuda))
    assert max_err(gx, gx_ref) < 1e-4
    m = smooth_flow_mask(flow, layout)
    assert max_err(gf.cpu().double() * m, gf_ref * m) < 2e-3 * max(1.0, gf_ref.abs().max().item())


def test_flow_warp_errors(cuda):
    x = torch.zeros(1, 4, 8, 8, device=cuda)
    with pytest.raises(ValueError):
        E.flow_warp(x, torch.zeros(1, 2, 8, 9, device=cuda))
    with pytest.raises(NotImplementedError):
        E.flow_warp(x, torch.zeros(1, 2, 8, 8, device=cuda), padding_mode="reflection")
    with pytest.raises(NotImplementedError):
        E.flow_warp(x.cpu(), torch.zeros(1, 2, 8, 8))


def test_flow_warp_size1_matches_reference_golden(cuda):
    d = np.load(os.path.join(GOLD, "flow_warp_size1.npz"))
    for tag, fn in (("n_row", E.flow_warp), ("n_col", E.flow_warp), ("m_px", E.flow_warp_nhw2)):
        x, flow, ref = (torch.from_numpy(d[f"{tag}_{k}"]) for k in ("x", "flow", "out"))
        xg = x.to(cuda).requires_grad_()
        fg = flow.to(cuda).requires_grad_()
        out = fn(xg, fg)
        assert max_err(out, ref) < 1e-5, tag
        out.sum().backward()
        assert torch.isfinite(fg.grad).all()


# ---------------------------------------------------------------------------------------------
# backwarp (BaseModel.get_backwarp, models/base_model.py:321-354; PWCNET.Decoder.backwarp,
# models/pwc_net.py:184-207)
# ---------------------------------------------------------------------------------------------
def test_backwarp_matches_reference_golden(cuda):
    d = np.load(os.path.join(GOLD, "backwarp.npz"))
    x, flow, ref, mref = (torch.from_numpy(d[k]) for k in ("x", "flow", "out", "mask"))
    out, mask = E.get_backwarp(x.to(cuda), flow.to(cuda))
    assert mask.shape == mref.shape and torch.equal(mask.cpu(), mref)
    assert max_err(out, ref) < 2e-4
    assert max_err(E.backwarp(x.to(cuda), flow.to(cuda)), ref) < 2e-4


@pytest.mark.parametrize("shape,cl,dtype", [((2, 3, 64, 48), False, torch.float32), ((1, 196, 2, 2), False, torch.float32),
                                            ((2, 32, 16, 16), False, torch.float32), ((1, 64, 17, 23), True, torch.float32),
                                            ((1, 3, 256, 256), False, torch.float32), ((2, 16, 12, 20), True, torch.bfloat16)])
def test_backwarp_vs_oracle(cuda, shape, cl, dtype):
    x, flow = warp_inputs(*shape, seed=21, sigma=1.5)
    x = x.to(dtype)
    ref, mref = O.backwarp(x.double(), flow.double())
    xg = x.to(cuda)
    xg = _cl(xg) if cl else xg
    out, mask = E.get_backwarp(xg, flow.to(cuda))
    assert out.dtype == dtype and out.shape == ref.shape
    # the threshold (warped ones > 0.999) is discontinuous: compare away from it
    py = torch.arange(shape[2]).view(1, -1, 1) + flow[:, 1].double() * shape[2] / (shape[2] - 1)
    px = torch.arange(shape[3]).view(1, 1, -1) + flow[:, 0].double() * shape[3] / (shape[3] - 1)
    edge = ((py.abs() < 2e-3) | ((py - (shape[2] - 1)).abs() < 2e-3) | (px.abs() < 2e-3) |
            ((px - (shape[3] - 1)).abs() < 2e-3)).unsqueeze(1)
    keep = (~edge).double()
    assert keep.mean() > 0.9
    assert ((mask.double().cpu() - mref) * keep).abs().max() == 0
    err = ((out.double().cpu() - ref) * keep).abs().max().item()
    assert err < (1e-4 if dtype == torch.float32 else BF16_REL * ref.pow(2).mean().sqrt().item())


def test_backwarp_backward(cuda):
    shape = (2, 6, 14, 18)
    x, flow = warp_inputs(*shape, seed=22, sigma=1.5)
    g = torch.randn(shape, generator=torch.Generator().manual_seed(23))
    xr = x.double().requires_grad_()
    fr = flow.double().requires_grad_()
    ref, _ = O.backwarp(xr, fr)
    gx_ref, gf_ref = torch.autograd.grad(ref, [xr, fr], g.double())
    xg = x.to(cuda).requires_grad_()
    fg = flow.to(cuda).requires_grad_()
    out, mask = E.get_backwarp(xg, fg)
    assert not mask.requires_grad
    gx, gf = torch.autograd.grad(out, [xg, fg], g.to(cuda))
    assert max_err(gx, gx_ref) < 1e-4
    sc = torch.tensor([shape[3] / (shape[3] - 1), shape[2] / (shape[2] - 1)]).view(1, 2, 1, 1)
    m = smooth_flow_mask(flow * sc)
    assert max_err(gf.cpu().double() * m, gf_ref * m) < 2e-3 * max(1.0, gf_ref.abs().max().item())


def test_backwarp_errors(cuda):
    x = torch.zeros(1, 3, 8, 8, device=cuda)
    with pytest.raises(ValueError):
        E.backwarp(x, torch.zeros(1, 2, 8, 9, device=cuda))
    with pytest.raises(NotImplementedError):
        E.backwarp(x.cpu(), torch.zeros(1, 2, 8, 8))
    with pytest.raises(L.EavsrError):
        E.backwarp(x[:, :, :1], torch.zeros(1, 2, 1, 8, device=cuda))


# ---------------------------------------------------------------------------------------------
# DCNv2 forward
# ---------------------------------------------------------------------------------------------
def _dcn_ref(x, off, mask, w, b, stride=1, padding=1, dilation=1, groups=1, dg=8):
    return O.modulated_deform_conv2d(x.double(), off.double(), mask.double(), w.double(),
                                     None if b is None else b.double(), stride, padding, dilation, groups, dg)


@pytest.mark.parametrize("dg", [8, 16, 1, 4])
@pytest.mark.parametrize("shape", [(1, 37, 53), (2, 16, 24), (1, 67, 120), (1, 11, 13)])
def test_dcn_tc_fp32(cuda, dg, shape):
    n, h, w = shape
    x, off, mask, wgt, bias = dcn_inputs(n, 64, h, w, 64, dg, seed=5)
    ref = _dcn_ref(x, off, mask, wgt, bias, dg=dg)
    xg = _cl(x.to(cuda))
    assert E.dcn_uses_tensor_cores(xg, wgt.to(cuda), 1, 1, 1, 1, dg)
    out = E.modulated_deform_conv2d(xg, off.to(cuda), mask.to(cuda), wgt.to(cuda), bias.to(cuda), 1, 1, 1, 1, dg)
    assert out.shape == ref.shape
    assert max_err(out, ref) < FP32_TOL
    assert max_err(out, ref) < 1e-4     # bf16x3 split keeps fp32-level accuracy


@pytest.mark.parametrize("dg", [8, 16])
def test_dcn_tc_bf16(cuda, dg):
    x, off, mask, wgt, bias = dcn_inputs(2, 64, 45, 61, 64, dg, seed=6)
    xb, wb, bb = x.bfloat16(), wgt.bfloat16(), bias.bfloat16()
    ref = _dcn_ref(xb, off, mask, wb, bb, dg=dg)
    out = E.modulated_deform_conv2d(_cl(xb.to(cuda)), off.to(cuda), mask.to(cuda), wb.to(cuda), bb.to(cuda),
                                    1, 1, 1, 1, dg)
    assert out.dtype == torch.bfloat16
    assert rel_err(out, ref) < BF16_REL


def test_dcn_static_weight_keeps_packed_image_and_tracks_updates(cuda):
    """static_weight=True skips the re-pack launch while the weight tensor is unchanged and re-packs
    after an in-place update (version counter) -- results always equal the uncached call."""
    x, off, mask, wgt, bias = dcn_inputs(1, 64, 40, 56, 64, 8, seed=16)
    xd, od, md = _cl(x.bfloat16().to(cuda)), off.to(cuda), mask.to(cuda)
    wd, bd = wgt.bfloat16().to(cuda), bias.bfloat16().to(cuda)

    def run(static):
        with torch.no_grad():
            return E.modulated_deform_conv2d(xd, od, md, wd, bd, 1, 1, 1, 1, 8, static_weight=static)
    base = run(False)
    n0 = L.launch_count()
    a = run(True)                       # packs
    n1 = L.launch_count()
    b = run(True)                       # reuses
    n2 = L.launch_count()
    assert torch.equal(a, base) and torch.equal(b, base)
    assert (n1 - n0) - (n2 - n1) == 1   # exactly the pack launch was saved
    with torch.no_grad():
        wd.mul_(0.5)                    # in-place update: must re-pack
    c = run(True)
    assert torch.equal(c, run(False))
    assert not torch.equal(c, base)


def test_dcn_tc_nchw_input_and_no_bias(cuda):
    x, off, mask, wgt, _ = dcn_inputs(1, 64, 21, 35, 64, 8, seed=7)
    ref = _dcn_ref(x, off, mask, wgt, None, dg=8)
    out = E.modulated_deform_conv2d(x.to(cuda), off.to(cuda), mask.to(cuda), wgt.to(cuda), None, 1, 1, 1, 1, 8)
    assert max_err(out, ref) < 1e-4


def test_dcn_zero_offset_equals_conv2d(cuda):
    x, _, _, wgt, bias = dcn_inputs(1, 64, 24, 40, 64, 8, seed=8)
    off = torch.zeros(1, 144, 24, 40)
    mask = torch.ones(1, 72, 24, 40)
    ref = torch.nn.functional.conv2d(x.double(), wgt.double(), bias.double(), padding=1)
    out = E.modulated_deform_conv2d(_cl(x.to(cuda)), off.to(cuda), mask.to(cuda), wgt.to(cuda), bias.to(cuda),
                                    1, 1, 1, 1, 8)
    assert max_err(out, ref) < 1e-4


@pytest.mark.parametrize("shape", [(1, 40, 72), (3, 67, 121), (2, 37, 52), (1, 270, 480)])
def test_dcn_tc_kernels_agree(cuda, shape):
    """warp-specialised tcgen05 kernel == first-generation tcgen05 kernel == generic SIMT kernel,
    including a multi-image batch with an odd width (scalar cp.async path) and the full bench size
    (persistent CTAs looping over several tiles, both TMEM accumulators in use)."""
    n, h, w = shape
    x, off, mask, wgt, bias = dcn_inputs(n, 64, h, w, 64, 8, seed=9)
    args = (_cl(x.to(cuda)), off.to(cuda), mask.to(cuda), wgt.to(cuda), bias.to(cuda), 1, 1, 1, 1, 8)
    a = _ModulatedDeformConv2dFn.apply(*args, 0)
    b = _ModulatedDeformConv2dFn.apply(*args, L.DCN_FORCE_V1)
    c = _ModulatedDeformConv2dFn.apply(*args, L.DCN_FORCE_GENERIC)
    assert max_err(a, c) < 1e-4
    assert max_err(b, c) < 1e-4
    # bf16: window kernel (bf16x2 blend / fp32 blend), warp-specialised L1 kernel, first-generation kernel
    xb, wb, bb = args[0].bfloat16(), wgt.to(cuda).bfloat16(), bias.to(cuda).bfloat16()
    run = lambda fl: _ModulatedDeformConv2dFn.apply(xb, *args[1:3], wb, bb, 1, 1, 1, 1, 8, fl)     # noqa: E731
    w16, w32, ws, v1 = run(0), run(L.DCN_BLEND_FP32), run(L.DCN_FORCE_WS), run(L.DCN_FORCE_V1)
    # fourth generation (TMA-staged offsets, one pixel per lane; taken by run(0) when w % 4 == 0) == third generation
    assert torch.equal(w16, run(L.DCN_FORCE_WIN1))
    assert torch.equal(w32, run(L.DCN_BLEND_FP32 | L.DCN_FORCE_WIN1))
    # fifth generation (A tile in tensor memory, taken by run(0)) == fourth (A tile in shared memory)
    assert torch.equal(w16, run(L.DCN_FORCE_WIN2))
    assert torch.equal(w32, run(L.DCN_BLEND_FP32 | L.DCN_FORCE_WIN2))
    assert torch.equal(ws, v1)              # same arithmetic, bit-identical
    assert torch.equal(w32, v1)             # the window only changes where the corners are read from
    ref = _dcn_ref(xb.cpu(), off, mask, wb.cpu(), bb.cpu(), dg=8)
    assert rel_err(w32, ref) < BF16_REL
    assert rel_err(w16, ref) < BF16_REL     # bf16x2 HFMA2 blend: a few more roundings, same tolerance
    print(f"dcn bf16 rel err: fp32-blend {rel_err(w32, ref):.2e}  bf16x2-blend {rel_err(w16, ref):.2e}")


@pytest.mark.parametrize("with_nan", [False, True])
@pytest.mark.parametrize("shape", [(1, 1, 4), (1, 3, 8), (2, 8, 16), (3, 9, 20), (1, 17, 132)])
def test_dcn_fifth_generation_edge_shapes_and_wild_offsets(cuda, shape, with_nan):
    """The TMEM-operand kernel (dcn_fwd_win3.cuh) on images smaller than / not a multiple of its 8 x 16 tile, a batch
    taken as a slice of a larger one (image stride != H*W*C), and offsets that are +-inf or astronomically large:
    bit-identical to the fourth generation and equal to the generic kernel (such samples contribute 0, as in mmcv's
    `h_im > -1 && h_im < height` test).  NaN offsets: the tcgen05 kernels propagate the NaN where mmcv's comparison
    would drop the sample -- a known, documented divergence; only the equality of the two generations
    is asserted for them."""
    n, h, w = shape
    x, off, mask, wgt, bias = dcn_inputs(n + 1, 64, h, w, 64, 8, seed=21)
    g = torch.Generator().manual_seed(3)
    r = torch.rand(off.shape, generator=g)
    if with_nan:
        off = torch.where(r < 0.0005, torch.full_like(off, float("nan")), off)          # ~7 % of the pixels
    off = torch.where((r >= 0.0005) & (r < 0.001), torch.full_like(off, float("inf")), off)
    off = torch.where((r >= 0.001) & (r < 0.003), torch.full_like(off, -3e38), off)
    off = torch.where((r >= 0.003) & (r < 0.005), torch.full_like(off, 2.0e9), off)
    xb = _cl(x.to(cuda).bfloat16())[1:]                       # images 1..n of a batch of n+1
    assert xb.stride(0) == h * w * 64 and xb.data_ptr() != xb.untyped_storage().data_ptr()
    args = (xb, off[1:].to(cuda).contiguous(), mask[1:].to(cuda).contiguous(), wgt.to(cuda).bfloat16(),
            bias.to(cuda).bfloat16(), 1, 1, 1, 1, 8)
    new, old = _ModulatedDeformConv2dFn.apply(*args, 0), _ModulatedDeformConv2dFn.apply(*args, L.DCN_FORCE_WIN2)
    assert torch.equal(torch.nan_to_num(new.float(), nan=123.0), torch.nan_to_num(old.float(), nan=123.0))
    if not with_nan:
        gen = _ModulatedDeformConv2dFn.apply(*args, L.DCN_FORCE_GENERIC).float()
        assert bool(torch.isfinite(gen).all()) and bool(torch.isfinite(new.float()).all())
        rms = gen.pow(2).mean().sqrt().item()
        assert (new.float() - gen).abs().max().item() <= 3e-2 * max(rms, 1e-3)


@pytest.mark.parametrize("shape", [(1, 40, 72), (2, 37, 53), (1, 270, 480)])
def test_dcn_window_kernel_deform_groups_16(cuda, shape):
    """BASELINE config 2 (deform_groups = 16): two 4-channel groups per 16-byte chunk in the window
    kernel.  With the fp32 blend it is bit-identical to the first-generation tcgen05 kernel; the default
    bf16x2 blend stays inside the bf16 tolerance of the fp64 oracle."""
    n, h, w = shape
    x, off, mask, wgt, bias = dcn_inputs(n, 64, h, w, 64, 16, seed=31)
    xb, wb, bb = _cl(x.bfloat16().to(cuda)), wgt.bfloat16().to(cuda), bias.bfloat16().to(cuda)
    run = lambda fl: _ModulatedDeformConv2dFn.apply(xb, off.to(cuda), mask.to(cuda), wb, bb, 1, 1, 1, 1, 16, fl)   # noqa: E731
    w16, w32, v1 = run(0), run(L.DCN_BLEND_FP32), run(L.DCN_FORCE_V1)
    assert torch.equal(w32, v1)
    ref = _dcn_ref(xb.cpu(), off, mask, wb.cpu(), bb.cpu(), dg=16)
    assert rel_err(w32, ref) < BF16_REL
    assert rel_err(w16, ref) < BF16_REL


@pytest.mark.parametrize("cfg", [
    dict(cin=16, cout=8, k=3, stride=1, padding=1, dilation=1, groups=1, dg=4),
    dict(cin=16, cout=8, k=3, stride=2, padding=2, dilation=2, groups=2, dg=4),
    dict(cin=6, cout=10, k=1, stride=1, padding=0, dilation=1, groups=1, dg=3),
    dict(cin=64, cout=64, k=3, stride=1, padding=1, dilation=1, groups=1, dg=64),
    dict(cin=128, cout=64, k=3, stride=1, padding=1, dilation=1, groups=1, dg=8),
])
@pytest.mark.parametrize("cl", [False, True])
def test_dcn_generic_fp32(cuda, cfg, cl):
    n, h, w = 2, 13, 17
    k, s, p, d = cfg["k"], cfg["stride"], cfg["padding"], cfg["dilation"]
    ho = (h + 2 * p - (d * (k - 1) + 1)) // s + 1
    wo = (w + 2 * p - (d * (k - 1) + 1)) // s + 1
    x, off, mask, wgt, bias = dcn_inputs(n, cfg["cin"], h, w, cfg["cout"], cfg["dg"], k=k, seed=10,
                                         groups=cfg["groups"], ho=ho, wo=wo)
    ref = _dcn_ref(x, off, mask, wgt, bias, s, p, d, cfg["groups"], cfg["dg"])
    xg = _cl(x.to(cuda)) if cl else x.to(cuda)
    out = E.modulated_deform_conv2d(xg, off.to(cuda), mask.to(cuda), wgt.to(cuda), bias.to(cuda), s, p, d,
                                    cfg["groups"], cfg["dg"])
    assert out.shape == ref.shape
    assert max_err(out, ref) < 1e-4


def test_dcn_module_and_errors(cuda):
    m = E.ModulatedDeformConv2d(64, 64, 3, padding=1, deform_groups=8).to(cuda)
    assert m.weight.shape == (64, 64, 3, 3) and m.bias.abs().sum().item() == 0
    x, off, mask, _, _ = dcn_inputs(1, 64, 12, 12, 64, 8, seed=11)
    out = m(x.to(cuda), off.to(cuda), mask.to(cuda))
    ref = _dcn_ref(x, off, mask, m.weight.detach().cpu(), m.bias.detach().cpu(), dg=8)
    assert max_err(out, ref) < 1e-4
    with pytest.raises(ValueError):
        E.modulated_deform_conv2d(x[0].to(cuda), off.to(cuda), mask.to(cuda), m.weight, m.bias, 1, 1, 1, 1, 8)
    with pytest.raises(ValueError):
        E.modulated_deform_conv2d(x.to(cuda), off[:, :18].to(cuda), mask.to(cuda), m.weight, m.bias, 1, 1, 1, 1, 8)
    with pytest.raises(NotImplementedError):
        E.modulated_deform_conv2d(x, off, mask, m.weight.cpu(), None, 1, 1, 1, 1, 8)


# ---------------------------------------------------------------------------------------------
# DCNv2 backward
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [
    dict(cin=64, cout=64, k=3, stride=1, padding=1, dilation=1, groups=1, dg=8, cl=True),
    dict(cin=64, cout=64, k=3, stride=1, padding=1, dilation=1, groups=1, dg=16, cl=True),
    dict(cin=16, cout=8, k=3, stride=2, padding=2, dilation=2, groups=2, dg=4, cl=False),
])
def test_dcn_backward_fp32(cuda, cfg):
    n, h, w = 2, 12, 15
    k, s, p, d = cfg["k"], cfg["stride"], cfg["padding"], cfg["dilation"]
    ho = (h + 2 * p - (d * (k - 1) + 1)) // s + 1
    wo = (w + 2 * p - (d * (k - 1) + 1)) // s + 1
    x, off, mask, wgt, bias = dcn_inputs(n, cfg["cin"], h, w, cfg["cout"], cfg["dg"], k=k, seed=12,
                                         groups=cfg["groups"], ho=ho, wo=wo, sigma=1.5)
    g = torch.randn(n, cfg["cout"], ho, wo, generator=torch.Generator().manual_seed(13))
    leaves = [t.double().requires_grad_() for t in (x, off, mask, wgt, bias)]
    ref = O.modulated_deform_conv2d(*leaves, s, p, d, cfg["groups"], cfg["dg"])
    grefs = torch.autograd.grad(ref, leaves, g.double())
    xg = _cl(x.to(cuda)) if cfg["cl"] else x.to(cuda)
    gl = [t.requires_grad_() for t in (xg, off.to(cuda), mask.to(cuda), wgt.to(cuda), bias.to(cuda))]
    out = E.modulated_deform_conv2d(*gl, s, p, d, cfg["groups"], cfg["dg"])
    grads = torch.autograd.grad(out, gl, g.to(cuda))
    for name, a, r in zip(("x", "offset", "mask", "weight", "bias"), grads, grefs):
        assert a.shape == r.shape, name
        assert max_err(a, r) < 1e-3 * max(1.0, r.abs().max().item()), name


def _dcn_bf16_grads(cuda, n, h, w, dg, flags, seed=21, sigma=1.5):
    x, off, mask, wgt, bias = dcn_inputs(n, 64, h, w, 64, dg, seed=seed, sigma=sigma)
    xb, wb, bb = x.bfloat16(), wgt.bfloat16(), bias.bfloat16()
    g = torch.randn(n, 64, h, w, generator=torch.Generator().manual_seed(seed + 1)).bfloat16()
    gl = [t.requires_grad_() for t in (_cl(xb.to(cuda)), off.to(cuda), mask.to(cuda), wb.to(cuda), bb.to(cuda))]
    out = _ModulatedDeformConv2dFn.apply(*gl, 1, 1, 1, 1, dg, flags)
    grads = torch.autograd.grad(out, gl, _cl(g.to(cuda)))
    return (xb, off, mask, wb, bb, g), grads


@pytest.mark.parametrize("flags", [0, L.DCN_BWD_GENERIC_DATA, L.DCN_BWD_GENERIC_WEIGHT])   # tcgen05 both / one half generic
@pytest.mark.parametrize("dg", [8, 16, 4, 1])
@pytest.mark.parametrize("shape", [(2, 19, 37), (1, 40, 72)])
def test_dcn_backward_bf16_tensor_cores(cuda, shape, dg, flags):
    """bf16 NHWC 64->64 backward on the tcgen05 kernels (csrc/dcn_bwd_tc.cu) against the fp64 oracle on
    the same bf16-rounded inputs; `flags` swaps one half at a time for the generic kernel."""
    n, h, w = shape
    inputs, grads = _dcn_bf16_grads(cuda, n, h, w, dg, flags)
    leaves = [t.double().requires_grad_() for t in inputs[:5]]
    ref = O.modulated_deform_conv2d(*leaves, 1, 1, 1, 1, dg)
    grefs = torch.autograd.grad(ref, leaves, inputs[5].double())
    tol = {"x": BF16_REL, "offset": 2e-3, "mask": 2e-3, "weight": BF16_REL, "bias": BF16_REL}
    for name, a, r in zip(("x", "offset", "mask", "weight", "bias"), grads, grefs):
        assert a.shape == r.shape, name
        assert rel_err(a, r) < tol[name], (name, rel_err(a, r))


@pytest.mark.parametrize("dg", [8, 16])
def test_dcn_backward_tensor_cores_match_generic_at_full_size(cuda, dg):
    """1x64x270x480 (BASELINE config 2/3 size): tcgen05 backward == generic backward."""
    _, tc = _dcn_bf16_grads(cuda, 1, 270, 480, dg, 0, seed=23, sigma=2.0)
    _, ge = _dcn_bf16_grads(cuda, 1, 270, 480, dg, L.DCN_BWD_GENERIC_DATA | L.DCN_BWD_GENERIC_WEIGHT, seed=23, sigma=2.0)
    tol = {"x": BF16_REL, "offset": 2e-3, "mask": 2e-3, "weight": BF16_REL, "bias": BF16_REL}
    for name, a, r in zip(("x", "offset", "mask", "weight", "bias"), tc, ge):
        assert rel_err(a, r) < tol[name], (name, rel_err(a, r))


# ---------------------------------------------------------------------------------------------
# correlation
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 32, 20, 32), (1, 196, 5, 8), (8, 64, 4, 4), (8, 196, 1, 1), (2, 96, 10, 16),
                                   (1, 7, 9, 13), (1, 16, 33, 70), (2, 20, 24, 72), (1, 32, 80, 128), (3, 9, 17, 44),
                                   (2, 64, 16, 16), (1, 5, 18, 16)])
@pytest.mark.parametrize("tf32", [False, True])
def test_correlation_forward_backward(cuda, shape, tf32, monkeypatch):
    """tf32=False: the default, exact fp32 kernels.  tf32=True: the opt-in tcgen05 banded GEMM (fp32 maps > 16x16 with
    w % 4 == 0; other shapes fall back to the exact kernels): kind::tf32 truncates the inputs to 10 mantissa bits, so
    the bound is relative to the output's scale."""
    from eavsr_b200 import ops
    monkeypatch.setattr(ops, "CORRELATION_TF32", tf32)
    g = torch.Generator().manual_seed(14)
    f1 = torch.randn(shape, generator=g)
    f2 = torch.randn(shape, generator=g)
    ref = O.correlation(f1.double(), f2.double())
    a = f1.to(cuda).requires_grad_()
    b = f2.to(cuda).requires_grad_()
    out = E.FunctionCorrelation(tenFirst=a, tenSecond=b)
    assert out.shape == ref.shape
    on_tc = tf32 and shape[2] * shape[3] > 256 and shape[3] % 4 == 0
    if on_tc:
        rms = ref.pow(2).mean().sqrt().item()
        assert max_err(out, ref) < 1.5e-2 * rms                   # worst element (2^-10 truncation of both inputs)
        assert (out.double().cpu() - ref).pow(2).mean().sqrt().item() < 3e-3 * rms
    else:
        assert max_err(out, ref) < 1e-5 * max(1.0, shape[1] ** 0.5)
    go = torch.randn(ref.shape, generator=g)
    g1r, g2r = O.correlation_backward(f1.double(), f2.double(), go.double())
    g1, g2 = torch.autograd.grad(out, [a, b], go.to(cuda))
    assert max_err(g1, g1r) < 1e-4
    assert max_err(g2, g2r) < 1e-4


def test_correlation_module_and_errors(cuda):
    m = E.ModuleCorrelation()
    f = torch.randn(1, 8, 6, 6)
    with pytest.raises(NotImplementedError):
        m(f, f)
    with pytest.raises(AssertionError):
        m(f.to(cuda).permute(0, 1, 3, 2), f.to(cuda))
    out = m(f.to(cuda), f.to(cuda))
    assert out.shape == (1, 81, 6, 6)


def test_launch_counter_counts_native_kernels(cuda):
    before = L.launch_count()
    x, flow = warp_inputs(1, 64, 8, 8)
    E.flow_warp(_cl(x.to(cuda)), flow.to(cuda))
    assert L.launch_count() == before + 1


# ---------------------------------------------------------------------------------------------
# size-independent properties at BASELINE's full sizes (no oracle needed)
# ---------------------------------------------------------------------------------------------
def test_full_size_properties(cuda):
    """1x64x270x480 (configs 2-4): DCNv2 is linear in x for fixed offsets / masks; a zero flow is the
    identity; the centre displacement of corr(f, f) is mean_c f^2 and corr is symmetric under swapping the
    arguments and negating the displacement (interior pixels)."""
    g = torch.Generator().manual_seed(71)
    h, w = 270, 480
    x1, off, mask, wgt, _ = dcn_inputs(1, 64, h, w, 64, 8, seed=72)
    x2 = torch.randn(1, 64, h, w, generator=g)
    a, b = 0.75, -1.25
    args = (off.to(cuda), mask.to(cuda), wgt.to(cuda), None, 1, 1, 1, 1, 8)
    with torch.no_grad():
        f = lambda t: E.modulated_deform_conv2d(_cl(t.to(cuda)), *args)       # noqa: E731
        lhs = f(a * x1 + b * x2)
        rhs = a * f(x1) + b * f(x2)
    assert max_err(lhs, rhs) < 2e-4 * max(1.0, rhs.abs().max().item())
    xb = _cl(x1.bfloat16().to(cuda))
    zero = torch.zeros(1, 2, h, w, device=cuda)
    assert torch.equal(E.flow_warp(xb, zero), xb)
    assert torch.equal(E.flow_warp(_cl(x1.to(cuda)), zero), _cl(x1.to(cuda)))
    fa = torch.randn(2, 32, 80, 128, generator=g).to(cuda)
    fb = torch.randn(2, 32, 80, 128, generator=g).to(cuda)
    c_aa = E.FunctionCorrelation(tenFirst=fa, tenSecond=fa)
    assert max_err(c_aa[:, 40], fa.pow(2).mean(1)) < 1e-5
    c_ab = E.FunctionCorrelation(tenFirst=fa, tenSecond=fb)
    c_ba = E.FunctionCorrelation(tenFirst=fb, tenSecond=fa)
    # c_ab[k(dy,dx)][y,x] == c_ba[k(-dy,-dx)][y+dy,x+dx]
    for dy, dx in ((1, -2), (-4, 4), (3, 0)):
        k, kn = (dy + 4) * 9 + (dx + 4), (-dy + 4) * 9 + (-dx + 4)
        lhs = c_ab[:, k, 8:-8, 8:-8]
        rhs = c_ba[:, kn, 8 + dy:c_ba.shape[2] - 8 + dy, 8 + dx:c_ba.shape[3] - 8 + dx]
        assert max_err(lhs, rhs) < 1e-5
