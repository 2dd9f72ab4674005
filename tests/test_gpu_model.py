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


# ---------------------------------------------------------------------------------------------
# BASELINE config 3: the 30-frame 270x480 clip, bf16, through the CUDA-graph path bench.py times
# ---------------------------------------------------------------------------------------------
def _residual(sr, lrs):
    """The part of the output the network computes: sr - bilinear x4 of the LR clip
    (models/eavsrp_model.py:350-364).  With seeded weights it is O(1e-2), so a max-abs bound on sr
    itself barely constrains it; the checks below are RELATIVE to it."""
    n, t, c, h, w = lrs.shape
    base = F.interpolate(lrs.reshape(n * t, c, h, w).float(), scale_factor=sr.shape[-1] // w, mode="bilinear",
                         align_corners=False)
    return sr.float() - base.view_as(sr)


def _feature_lists(net, lrs):
    """Run `net` and capture the four propagation branches' outputs (the recurrent state that carries the
    alignment error over 30 steps), as (n, t, 64, h, w) fp32 tensors."""
    got = {}
    orig = net._propagate

    def spy(feats, flows, branch):
        feats = orig(feats, flows, branch)
        got[branch] = torch.stack([f.float() for f in feats[branch]], 1)
        return feats

    net._propagate = spy
    try:
        with torch.no_grad():
            sr = net(lrs)
    finally:
        del net._propagate
    return sr, got


def test_config3_bf16_graph_path_vs_fp32_and_cpu_oracle(cuda):
    """(i) bf16 CUDA-graph replay of the full 30x270(->272)x480 clip == eager bf16 run (to rounding);
    (ii) against the fp32 path of the same kernels (golden-pinned above): SR max-abs <= 1e-2, PSNR delta
    <= 0.01 dB, learned residual within 3 % relative RMS, every branch feature list within 3 % relative RMS
    with a stated worst-element bound; (iii) a T=6 272x480 clip against the CPU oracle (fp32 <= 1e-3 on SR
    and <= 2 % on the residual; bf16 <= 1e-2 / 3 %)."""
    import copy
    t, h, w = 30, 270, 480
    net = EAVSRP(4).eval()
    seeded_parameters(net)
    net = net.to(cuda).prepare(torch.bfloat16)
    lrs = pad_clip(clip_inputs(1, t, h, w, seed=1234).to(cuda))
    static_in = lrs.clone()
    with torch.no_grad():
        eager = net(static_in)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            net(static_in)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = net(static_in)
        static_in.copy_(pad_clip(clip_inputs(1, t, h, w, seed=1).to(cuda)))     # the graph must read its input buffer
        graph.replay()
        assert not torch.equal(out, eager)
        static_in.copy_(lrs)
        graph.replay()
        torch.cuda.synchronize()
        sr16 = out.float().clone()
    # (i) the channel-attention sums are accumulated with floating-point atomics (order varies run to run), so two
    # runs of the same clip agree to rounding, not bit for bit: compare on the learned residual
    eager = eager.float()
    r_g, r_e = _residual(sr16, lrs), _residual(eager, lrs)
    assert (sr16 - eager).abs().max().item() < 1e-3
    assert (r_g - r_e).pow(2).mean().sqrt().item() <= 0.01 * r_e.pow(2).mean().sqrt().item()
    del graph, out, eager, r_g, r_e

    sr16b, f16 = _feature_lists(net, lrs)
    assert (sr16b.float() - sr16).abs().max().item() < 1e-3
    net32 = copy.deepcopy(net).float().prepare(torch.float32)      # the bf16-rounded weights, evaluated in fp32
    sr32, f32 = _feature_lists(net32, lrs)
    sr32 = sr32.float()
    assert (sr16 - sr32).abs().max().item() < 1e-2                                                     # north_star
    crop = lambda x: x[..., : 4 * h, : 4 * w]       # noqa: E731
    hr = F.interpolate(lrs[0, :, :, :h, :w], scale_factor=4, mode="bicubic", align_corners=False).clamp(0, 1)[None]
    assert abs(psnr(visuals(crop(sr16)), visuals(hr)) - psnr(visuals(crop(sr32)), visuals(hr))) <= 0.01
    r16, r32 = _residual(sr16, lrs), _residual(sr32, lrs)
    rms = r32.pow(2).mean().sqrt().item()
    assert rms > 1e-3                                               # the residual is not trivially zero
    assert (r16 - r32).pow(2).mean().sqrt().item() <= 0.03 * rms
    assert (r16 - r32).abs().max().item() <= 0.5 * rms + 2e-3      # worst element: bf16 output rounding of a [0,1] frame
    for b in f32:
        ref_rms = f32[b].pow(2).mean().sqrt().item()
        d = f16[b] - f32[b]
        assert d.pow(2).mean().sqrt().item() <= 0.03 * ref_rms, b
        assert d.abs().max().item() <= 0.25 * f32[b].abs().max().item(), b
        # error growth over the 30 recurrent steps stays bounded: the last-visited frame is no worse than 3x the first
        per_t = d.pow(2).mean((0, 2, 3, 4)).sqrt()
        assert per_t.max().item() <= 3.0 * per_t[per_t > 0].min().item() + 1e-3 * ref_rms, b
    del f16, f32, sr16b

    # (iii) T=6 at the full frame size against the CPU oracle (restatement kernels, ~1 min on 8 cores)
    lr6 = lrs[:, :6].clone()
    sd = {k: v.detach().float().cpu() for k, v in net32.state_dict().items()}
    with torch.no_grad():
        ref = eavsrp_cpu.eavsrp_forward(sd, lr6.cpu(), 4, "aten")
        o32 = net32(lr6).float().cpu()
        o16 = net(lr6).float().cpu()
    assert (o32 - ref).abs().max().item() < 1e-3
    assert (o16 - ref).abs().max().item() < 1e-2
    rr = _residual(ref, lr6.cpu())
    rrms = rr.pow(2).mean().sqrt().item()
    assert (_residual(o32, lr6.cpu()) - rr).pow(2).mean().sqrt().item() <= 0.02 * rrms
    assert (_residual(o16, lr6.cpu()) - rr).pow(2).mean().sqrt().item() <= 0.03 * rrms


# ---------------------------------------------------------------------------------------------
# PWC-Net (train-time caller of the cost volume and of backwarp) against the reference's own modules
# ---------------------------------------------------------------------------------------------
def test_pwcnet_estimate_and_get_backwarp_match_reference_golden(cuda):
    from eavsr_b200 import pwc
    g = np.load(os.path.join(GOLD, "pwc_net.npz"))
    t = {k: torch.from_numpy(g[k]).to(cuda) for k in ("first", "second", "flow_net", "lr", "hr", "flow_est", "out", "mask")}
    net = pwc.PWCNET().eval()
    seeded_parameters(net)
    net = net.to(cuda)
    with torch.no_grad():
        flow = net(t["first"], t["second"])
        est = pwc.estimate(t["lr"], F.interpolate(t["hr"], scale_factor=0.5, mode="bilinear", align_corners=True), net)
        out, mask = pwc.get_backwarp(t["lr"], t["hr"], net, scale=2)
    assert (flow - t["flow_net"]).abs().max().item() < 1e-3 * max(1.0, t["flow_net"].abs().max().item())
    assert (est - t["flow_est"]).abs().max().item() < 2e-3 * max(1.0, t["flow_est"].abs().max().item())
    agree = (mask == t["mask"])
    assert agree.float().mean().item() > 0.995          # the 0.999 threshold is discontinuous in the flow
    assert ((out - t["out"]) * agree).abs().max().item() < 2e-3
