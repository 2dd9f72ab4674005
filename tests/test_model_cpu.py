"""CPU tests of the host-side callers (eavsr_b200.model): state-dict compatibility with the
reference and -- with the three CUDA operators swapped for the CPU oracle by the TEST (the product
never imports oracle/) -- output parity with the golden vectors the reference EAVSRP produced."""
import os

import numpy as np
import pytest
import torch

import eavsr_b200.model as M
from oracle import alignment as O

from helpers import clip_inputs, seeded_parameters

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture
def oracle_ops(monkeypatch):
    monkeypatch.setattr(M, "flow_warp", lambda x, f, padding_mode="zeros": O.flow_warp(x, f.to(x.dtype), "n2hw", padding_mode))
    monkeypatch.setattr(M, "flow_warp_nhw2", lambda x, f, padding_mode="zeros": O.flow_warp(x, f.to(x.dtype), "nhw2", padding_mode))
    monkeypatch.setattr(M, "modulated_deform_conv2d",
                        lambda x, off, m, w, b, s, p, d, g, dg, static_weight=False: O.modulated_deform_conv2d(x, off, m, w, b, s, p, d, g, dg))


@pytest.mark.parametrize("scale,t", [(4, 4), (2, 3)])
def test_state_dict_matches_reference(scale, t):
    g = np.load(os.path.join(GOLD, f"eavsrp_x{scale}.npz"))
    net = M.EAVSRP(scale)
    sd = net.state_dict()
    assert sorted(sd.keys()) == list(g["keys"])
    assert [str(tuple(sd[k].shape)) for k in sorted(sd.keys())] == list(g["key_shapes"])
    assert sum(p.numel() for p in net.parameters()) == int(g["nparams"])
    assert not any(p.requires_grad for p in net.spynet.parameters())


@pytest.mark.parametrize("scale,t", [(4, 4), (2, 3)])
def test_forward_matches_reference_golden(oracle_ops, scale, t):
    g = np.load(os.path.join(GOLD, f"eavsrp_x{scale}.npz"))
    net = M.EAVSRP(scale).eval()
    seeded_parameters(net)
    lrs = clip_inputs(1, t, 64, 64, seed=107)
    with torch.no_grad():
        sr = net(lrs)
    assert list(sr.shape) == list(g["shape"])
    assert (sr[..., ::4, ::4] - torch.from_numpy(g["sr_sub"])).abs().max() < 2e-4
    assert (sr[..., 1::8, 2::8] - torch.from_numpy(g["sr_sub2"])).abs().max() < 2e-4
    assert abs(sr.mean().item() - float(g["sr_mean"])) < 1e-5


def test_multi_adstn_matches_reference_golden(oracle_ops):
    g = np.load(os.path.join(GOLD, "multi_adstn.npz"))
    m = M.MultiAdSTN(64, 8).eval()
    seeded_parameters(m)
    t = lambda k: torch.from_numpy(g[k])    # noqa: E731
    with torch.no_grad():
        out = m([t("nbr0"), t("nbr1"), t("nbr2")], [t("ref0"), t("ref1"), t("ref2")], t("prop"), t("flow"))
    assert (out - t("out")).abs().max() < 1e-4


def test_affine_offsets_match_reference_golden():
    g = np.load(os.path.join(GOLD, "adapt_block_offset.npz"))
    off = M._affine_offsets(torch.from_numpy(g["transform"]), torch.from_numpy(g["translation"]), 8)
    assert (off - torch.from_numpy(g["offset"])).abs().max() < 1e-5


def test_rejects_sizes_the_reference_cannot_run():
    net = M.EAVSRP(4)
    with pytest.raises(ValueError, match="divisible by 4"):
        net(torch.zeros(1, 3, 3, 66, 64))
    with pytest.raises(AssertionError):
        net(torch.zeros(1, 3, 3, 32, 64))
    assert M.pad_clip(torch.zeros(1, 2, 3, 270, 480)).shape == (1, 2, 3, 272, 480)
