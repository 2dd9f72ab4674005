"""GPU integration parity: eavsr_b200.model.EAVSRP (CUDA alignment kernels) against the golden
vectors the reference EAVSRP produced and against the functional CPU oracle, with identical seeded
weights and clips.  Tolerances from BASELINE.json: max-abs <= 1e-3 (fp32) / <= 1e-2 (bf16) on [0,1]
frames and <= 0.01 dB PSNR delta (reference's calc_psnr on clamp(x*255).round() visuals)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from eavsr_b200.model import EAVSRP, pad_clip
from eavsr_b200.synthetic import clip_inputs, seeded_parameters
from oracle import eavsrp_cpu

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def visuals(x):                      # models/base_model.py:146-150
    return torch.clamp(x.float() * 255, 0, 255).round()


def psnr(sr, hr, rng=255.0):         # util/util.py:302-320
    return (-10 * torch.log10(torch.pow((sr - hr) / rng, 2).mean())).item()


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("scale,t", [(4, 4), (2, 3)])
def test_fp32_matches_reference_golden(cuda, scale, t):
    g = np.load(os.path.join(GOLD, f"eavsrp_x{scale}.npz"))
    net = EAVSRP(scale).eval()
    seeded_parameters(net)
    net = net.to(cuda).prepare(torch.float32)
    lrs = clip_inputs(1, t, 64, 64, seed=107)
    with torch.no_grad():
        sr = net(lrs.to(cuda)).cpu()
    assert list(sr.shape) == list(g["shape"])
    assert (sr[..., ::4, ::4] - torch.from_numpy(g["sr_sub"])).abs().max() < 1e-3
    assert (sr[..., 1::8, 2::8] - torch.from_numpy(g["sr_sub2"])).abs().max() < 1e-3


def test_bf16_and_psnr_delta_vs_cpu_oracle(cuda):
    t, h, w = 5, 68, 96
    net = EAVSRP(4).eval()
    seeded_parameters(net)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    lrs = clip_inputs(1, t, h, w, seed=11)
    hr = F.interpolate(lrs[0], scale_factor=4, mode="bicubic", align_corners=False).clamp(0, 1).unsqueeze(0)
    with torch.no_grad():
        ref = eavsrp_cpu.eavsrp_forward(sd, lrs, 4, "restatement")
        out32 = net.to(cuda).prepare(torch.float32)(lrs.to(cuda)).float().cpu()
        out16 = net.prepare(torch.bfloat16)(lrs.to(cuda)).float().cpu()
    assert (out32 - ref).abs().max() < 1e-3
    assert (out16 - ref).abs().max() < 1e-2
    p_ref = psnr(visuals(ref), visuals(hr))
    assert abs(psnr(visuals(out32), visuals(hr)) - p_ref) <= 0.01
    assert abs(psnr(visuals(out16), visuals(hr)) - p_ref) <= 0.01


def test_padded_270_clip_runs_and_crops(cuda):
    net = EAVSRP(4).eval()
    seeded_parameters(net)
    net = net.to(cuda).prepare(torch.bfloat16)
    lrs = clip_inputs(1, 3, 70, 90, seed=3)           # 70, 90 are not multiples of 4, like 270
    with torch.no_grad():
        sr = net(pad_clip(lrs.to(cuda)))[..., :280, :360]
    assert sr.shape == (1, 3, 3, 280, 360) and torch.isfinite(sr.float()).all()


def test_state_dict_roundtrip_strict(cuda):
    a = EAVSRP(4)
    seeded_parameters(a)
    b = EAVSRP(4)
    b.load_state_dict(a.state_dict(), strict=True)
