"""CPU tests: the oracle against the golden fixtures produced by the reference itself
(tests/golden/make_golden.py) and against live torchvision / ATen stand-ins."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import alignment as O

from helpers import dcn_inputs, dcn_offset_grad_mask, smooth_flow_mask, warp_inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files if d[k].dtype.kind == "f"}


@pytest.mark.parametrize("tag,layout,pad", [("networks_zeros", "n2hw", "zeros"), ("model_zeros", "nhw2", "zeros"),
                                            ("model_border", "nhw2", "border")])
def test_flow_warp_matches_reference(tag, layout, pad):
    g = gold("flow_warp_" + tag)
    x = g["x"].double().requires_grad_()
    flow = g["flow"].double().requires_grad_()
    out = O.flow_warp(x, flow, layout, pad)
    assert (out - g["out"]).abs().max() < 2e-4   # reference ran in fp32 incl. its normalise/un-normalise
    gx, gf = torch.autograd.grad(out, [x, flow], g["g"].double())
    assert (gx - g["gx"]).abs().max() < 1e-4
    m = smooth_flow_mask(g["flow"], layout)
    assert m.mean() > 0.9
    assert ((gf - g["gflow"]) * m).abs().max() < 2e-3 * max(1.0, g["gflow"].abs().max().item())


def test_flow_warp_size1_matches_reference():
    g = gold("flow_warp_size1")
    for tag, layout in (("n_row", "n2hw"), ("n_col", "n2hw"), ("m_px", "nhw2")):
        out = O.flow_warp(g[tag + "_x"].double(), g[tag + "_flow"].double(), layout)
        assert (out - g[tag + "_out"]).abs().max() < 1e-5, tag


def test_flow_warp_inputs_are_the_seeded_ones():
    g = gold("flow_warp_networks_zeros")
    x, flow = warp_inputs(2, 6, 11, 14, seed=100)
    assert torch.equal(x, g["x"]) and torch.equal(flow, g["flow"])


def test_backwarp_matches_reference():
    g = gold("backwarp")
    out, mask = O.backwarp(g["x"].double(), g["flow"].double())
    assert torch.equal(mask.float(), g["mask"])
    assert (out - g["out"]).abs().max() < 2e-4


@pytest.mark.parametrize("dg", [8, 16])
def test_dcn_matches_reference_call_site(dg):
    g = gold(f"dcn_dg{dg}")
    x, off, mask, w, b = dcn_inputs(1, 64, 9, 11, 64, dg, seed=103)
    leaves = [t.double().requires_grad_() for t in (x, off, mask, w, b)]
    out = O.modulated_deform_conv2d(*leaves, 1, 1, 1, 1, dg)
    assert (out - g["out"]).abs().max() < 1e-4
    grads = torch.autograd.grad(out, leaves, g["g"].double())
    for name, a in zip(("gx", "goffset", "gmask", "gweight", "gbias"), grads):
        m = dcn_offset_grad_mask(off) if name == "goffset" else 1.0
        assert ((a - g[name]) * m).abs().max() < 1e-3 * max(1.0, g[name].abs().max().item()), name


def test_affine_offsets_match_adapt_block_offset():
    g = gold("adapt_block_offset")
    off = O.affine_offsets(g["transform"].double(), g["translation"].double(), 8)
    assert (off - g["offset"]).abs().max() < 1e-5


@pytest.mark.parametrize("tag", ["a", "b"])
def test_correlation_matches_reference_kernels(tag):
    g = gold("correlation_" + tag)
    out = O.correlation(g["first"].double(), g["second"].double())
    assert (out - g["out"]).abs().max() < 1e-5
    g1, g2 = O.correlation_backward(g["first"].double(), g["second"].double(), g["gout"].double())
    assert (g1 - g["gfirst"]).abs().max() < 1e-5
    assert (g2 - g["gsecond"]).abs().max() < 1e-5


def test_correlation_backward_is_the_adjoint():
    gen = torch.Generator().manual_seed(0)
    f1 = torch.randn(2, 5, 6, 7, generator=gen, dtype=torch.float64, requires_grad=True)
    f2 = torch.randn(2, 5, 6, 7, generator=gen, dtype=torch.float64, requires_grad=True)
    out = O.correlation(f1, f2)
    go = torch.randn(out.shape, generator=gen, dtype=torch.float64)
    a = torch.autograd.grad(out, [f1, f2], go)
    b = O.correlation_backward(f1, f2, go)
    assert (a[0] - b[0]).abs().max() < 1e-12 and (a[1] - b[1]).abs().max() < 1e-12


def test_dcn_against_torchvision_live():
    torchvision = pytest.importorskip("torchvision")
    x, off, mask, w, b = dcn_inputs(2, 16, 13, 17, 8, 4, seed=1, groups=2, ho=7, wo=9)
    ref = torchvision.ops.deform_conv2d(x.double(), off.double(), w.double(), b.double(), stride=2, padding=2,
                                        dilation=2, mask=mask.double())
    out = O.modulated_deform_conv2d(x.double(), off.double(), mask.double(), w.double(), b.double(), 2, 2, 2, 2, 4)
    assert (out - ref).abs().max() < 1e-12


def test_dcn_zero_offset_is_conv2d():
    x, _, _, w, b = dcn_inputs(1, 8, 9, 9, 4, 2, seed=2)
    off = torch.zeros(1, 36, 9, 9, dtype=torch.float64)
    mask = torch.ones(1, 18, 9, 9, dtype=torch.float64)
    out = O.modulated_deform_conv2d(x.double(), off, mask, w.double(), b.double(), 1, 1, 1, 1, 2)
    assert (out - F.conv2d(x.double(), w.double(), b.double(), padding=1)).abs().max() < 1e-12


@pytest.mark.parametrize("pad", ["zeros", "border"])
def test_flow_warp_against_grid_sample_live(pad):
    x, flow = warp_inputs(2, 3, 10, 12, seed=3)
    x, flow = x.double(), flow.double()
    h, w = 10, 12
    gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    gf = torch.stack((gx, gy), 2).double() + flow.permute(0, 2, 3, 1)
    grid = torch.stack((2 * gf[..., 0] / (w - 1) - 1, 2 * gf[..., 1] / (h - 1) - 1), 3)
    ref = F.grid_sample(x, grid, mode="bilinear", padding_mode=pad, align_corners=True)
    assert (O.flow_warp(x, flow, "n2hw", pad) - ref).abs().max() < 1e-9


def test_flow_warp_size_mismatch_raises():
    with pytest.raises(ValueError):
        O.flow_warp(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 5))
