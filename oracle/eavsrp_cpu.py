"""Functional CPU restatement of the reference EAVSR+ generator (x4 and x2).

TEST / BENCH INFRASTRUCTURE -- see oracle/__init__.py.  Written against a plain state dict with
torch.nn.functional so that it shares no code with eavsr_b200/model.py; pinned by the golden
vectors the reference itself produced (tests/golden/eavsrp_x{4,2}.npz).  Follows
models/eavsrp_model.py:179-364 (forward / compute_flow / propagate / upsample), :433-523 (SPyNet),
models/networks.py:280-348 (AdaptBlock*), :522-552 (encoder), :597-631 (MultiAdSTN).

`ops` selects the alignment operators:
  * "restatement": oracle.alignment (explicit index arithmetic; any dtype, e.g. fp64)
  * "aten":        F.grid_sample + torchvision.ops.deform_conv2d -- what the reference's
                   `--gpu_ids -1` path executes on the host (mmcv stand-in), used for the timed
                   CPU baseline.
"""
from __future__ import annotations

import time

import torch
import torch.nn.functional as F

from . import alignment as A
from . import cpu_reference as R

BRANCHES = ("backward_1", "forward_1", "backward_2", "forward_2")


class Ops:
    def __init__(self, kind="restatement"):
        self.kind = kind

    def warp(self, x, flow_n2hw, pad="zeros"):
        if self.kind == "aten":
            return R.ref_flow_warp(x, flow_n2hw.permute(0, 2, 3, 1), pad)
        return A.flow_warp(x, flow_n2hw, "n2hw", pad)

    def dcn(self, x, off, mask, w, b, dg):
        if self.kind == "aten":
            return R.ref_dcn(x, off, mask, w, b)
        return A.modulated_deform_conv2d(x, off, mask, w, b, 1, 1, 1, 1, dg)


def _conv(sd, p, x, pad, groups=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad, groups=groups)


def _rca_group(sd, p, x, nb):
    y = x
    for i in range(nb):
        q = f"{p}.rg.{i}"
        r = _conv(sd, q + ".res.2", F.relu(_conv(sd, q + ".res.0", y, 1)), 1)
        s = r.mean((2, 3), keepdim=True)
        s = torch.sigmoid(_conv(sd, q + ".ca.conv_du.2", F.relu(_conv(sd, q + ".ca.conv_du.0", s, 0)), 0))
        y = r * s + y
    return _conv(sd, f"{p}.rg.{nb}", y, 1) + x


def _res_stack(sd, p, x, nb):
    return _rca_group(sd, p + ".main.2", F.leaky_relu(_conv(sd, p + ".main.0", x, 1), 0.1), nb)


def _mix(sd, p, a, b):
    y = F.leaky_relu(_conv(sd, p + ".concat.0", torch.cat([a, b], 1), 1, groups=a.shape[1] * 2), 0.2)
    return F.leaky_relu(_conv(sd, p + ".concat2.0", y, 1, groups=a.shape[1]), 0.2)


def _adapt(sd, p, a, b, D, k):
    f = _mix(sd, p, a, b)
    off = A.affine_offsets(_conv(sd, p + ".transform_matrix_conv", f, k // 2),
                           _conv(sd, p + ".translation_conv", f, k // 2), D)
    return f, off


def _resize(flow, s):
    return F.interpolate(flow, scale_factor=s, mode="bilinear", align_corners=True) * s


def multi_adstn(sd, p, ops, nbr, ref, prop, flow, dg=8):
    f4, f2 = _resize(flow, 0.25), _resize(flow, 0.5)
    res = lambda lvl, w_, r_: _conv(sd, f"{p}.trans_l{lvl}.conv_first",            # noqa: E731
                                    _adapt(sd, f"{p}.flow_l{lvl}", w_, r_, 1, 3)[1], 1)
    p1u = _resize(res(3, ops.warp(nbr[2], f4), ref[2]), 2)
    p2 = res(2, ops.warp(nbr[1], f2 + p1u), ref[1])
    p2u = _resize(p2 + p1u, 2)
    p3 = res(1, ops.warp(nbr[0], flow + p2u), ref[0])
    flow = p3 + p2u + flow
    nbr_w, feat = ops.warp(nbr[0], flow), ops.warp(prop, flow)
    f, off = _adapt(sd, p + ".adastn", nbr_w, ref[0], dg, 5)
    mask = torch.sigmoid(_conv(sd, p + ".adastn.mask_conv", f, 2))
    return ops.dcn(feat, off, mask, sd[p + ".weight"], sd[p + ".bias"], dg)


def spynet(sd, ops, ref, supp):
    h, w = ref.shape[2:]
    hu, wu = -(-h // 32) * 32, -(-w // 32) * 32
    norm = lambda x: (F.interpolate(x, size=(hu, wu), mode="bilinear", align_corners=False)   # noqa: E731
                      - sd["spynet.mean"]) / sd["spynet.std"]
    ref, supp = [norm(ref)], [norm(supp)]
    for _ in range(5):
        ref.append(F.avg_pool2d(ref[-1], 2, 2, count_include_pad=False))
        supp.append(F.avg_pool2d(supp[-1], 2, 2, count_include_pad=False))
    ref, supp = ref[::-1], supp[::-1]
    flow = ref[0].new_zeros(ref[0].shape[0], 2, hu // 32, wu // 32)
    for lvl in range(6):
        up = flow if lvl == 0 else _resize(flow, 2)
        x = torch.cat([ref[lvl], ops.warp(supp[lvl], up, "border"), up], 1)
        for i in range(5):
            x = _conv(sd, f"spynet.basic_module.{lvl}.basic_module.{i}.conv", x, 3)
            if i < 4:
                x = F.relu(x)
        flow = up + x
    flow = F.interpolate(flow, size=(h, w), mode="bilinear", align_corners=False)
    return flow * flow.new_tensor([w / wu, h / hu]).view(1, 2, 1, 1)


def eavsrp_forward(sd, lrs, scale=4, ops_kind="restatement", dg=8, nb=30):
    ops = Ops(ops_kind)
    n, t, c, h, w = lrs.shape
    a, b = lrs[:, :-1].reshape(-1, c, h, w), lrs[:, 1:].reshape(-1, c, h, w)
    flows = {"backward": spynet(sd, ops, a, b).view(n, t - 1, 2, h, w),
             "forward": spynet(sd, ops, b, a).view(n, t - 1, 2, h, w)}
    x = (lrs.reshape(-1, c, h, w) - sd["encoder.mean"]) / sd["encoder.std"]
    for name in ("conv1_1", "conv1_2", "conv2_1", "conv2_2"):
        x = F.relu(_conv(sd, "encoder.model." + name, x, 1))
    f1 = _conv(sd, "encoder.tail", _conv(sd, "encoder.model.conv3_1", x, 1), 1)
    pyr = [f1] + [F.interpolate(f1, scale_factor=s, mode="bilinear", align_corners=False) for s in (0.5, 0.25)]
    pyr = [p.view(n, t, *p.shape[1:]) for p in pyr]
    level = lambda j: [p[:, j] for p in pyr]        # noqa: E731
    done = {}
    for bi, br in enumerate(BRANCHES):
        back = br.startswith("backward")
        fl = flows["backward" if back else "forward"]
        order = list(range(t - 1, -1, -1)) if back else list(range(t))
        step = 1 if back else -1
        prop = f1.new_zeros(n, 64, h, w)
        outs, prev = [], None
        for i, idx in enumerate(order):
            cur = pyr[0][:, idx]
            if i > 0:
                f_1 = fl[:, idx if back else idx - 1]
                c1 = multi_adstn(sd, "deform_align." + br, ops, level(idx + step), level(idx), prop, f_1, dg)
                c2 = torch.zeros_like(c1)
                if i > 1:
                    f_2 = f_1 + ops.warp(prev, f_1)
                    c2 = multi_adstn(sd, "deform_align." + br, ops, level(idx + 2 * step), level(idx), outs[-2], f_2, dg)
                prop = _conv(sd, "fusion." + br, torch.cat([c1, cur, c2], 1), 0)
                prev = f_1
            feat = torch.cat([cur] + [done[k][idx] for k in BRANCHES[:bi]] + [prop], 1)
            prop = prop + _res_stack(sd, "backbone." + br, feat, nb)
            outs.append(prop)
        done[br] = outs[::-1] if back else outs
    frames = []
    for i in range(t):
        y = _res_stack(sd, "reconstruction", torch.cat([pyr[0][:, i]] + [done[k][i] for k in BRANCHES], 1), 5)
        y = F.leaky_relu(F.pixel_shuffle(_conv(sd, "upsample1.0", y, 1), 2), 0.1)
        if scale == 4:
            y = F.leaky_relu(F.pixel_shuffle(_conv(sd, "upsample2.0", y, 1), 2), 0.1)
        y = _conv(sd, "conv_last", F.leaky_relu(_conv(sd, "conv_hr", y, 1), 0.1), 1)
        frames.append(y + F.interpolate(lrs[:, i], scale_factor=scale, mode="bilinear", align_corners=False))
    return torch.stack(frames, 1)


def _seeded_state_dict():
    from eavsr_b200.model import EAVSRP
    from eavsr_b200.synthetic import seeded_parameters
    net = EAVSRP(4)
    seeded_parameters(net)
    return {k: v.float() for k, v in net.state_dict().items()}


def time_clip_forward(t_s, h, w, state_dict=None, seed=1234):
    """Seconds for the complete x4 forward of one `t_s`-frame h x w clip through the reference's CPU operators
    (ATen grid_sample + torchvision CPU DCNv2 = what `--gpu_ids -1` executes), fp32, all host threads."""
    from eavsr_b200.synthetic import clip_inputs
    sd = state_dict or _seeded_state_dict()
    lrs = clip_inputs(1, t_s, h, w, seed=seed)
    with torch.no_grad():
        t0 = time.time()
        eavsrp_forward(sd, lrs, 4, "aten")
        return time.time() - t0


def plan_sample(budget_s, h=272, w=480, state_dict=None, frames=(6, 4, 3)):
    """Pick the largest sample that fits `budget_s` on this host: a tiny probe clip gives seconds per LR
    pixel-frame; prefer the stated geometry (h x w) with as many frames as fit, fall back to a reduced frame
    size only when even 3 frames at h x w do not fit.  Returns (t_s, hs, ws, estimated seconds)."""
    time_clip_forward(3, 64, 64, state_dict)                   # pays the lazy initialisations (~20 s of oneDNN set-up)
    probe = time_clip_forward(3, 64, 96, state_dict)
    per_pxf = probe / (3 * 64 * 96) * 0.8                      # large frames thread better than the probe (measured)
    for t_s in frames:
        est = per_pxf * t_s * h * w
        if est <= budget_s:
            return t_s, h, w, est
    hs, ws = h, w
    while per_pxf * 3 * hs * ws > budget_s and hs > 64:
        hs, ws = (hs // 2 + 3) // 4 * 4, (ws // 2 + 3) // 4 * 4
    return 3, hs, ws, per_pxf * 3 * hs * ws


def time_model_sample(budget_s=30.0, h=272, w=480, t_full=30, state_dict=None, steps=1, frames=(6, 4, 3)):
    """CPU baseline for the full-clip metric on a BOUNDED sample: `steps` complete x4 forwards of a short clip
    at the stated frame size (272x480) when the host is fast enough for `budget_s` per step.
    Returns (LR frames/s at h x w, description, seconds per step list, (t_s, hs, ws))."""
    sd = state_dict or _seeded_state_dict()
    t_s, hs, ws, _ = plan_sample(budget_s, h, w, sd, frames)
    secs = [time_clip_forward(t_s, hs, ws, sd, seed=1234 + i) for i in range(max(1, steps))]
    el = sum(secs) / len(secs)
    # a T=30 clip does (2T-3)/T = 1.9 alignments per frame against 1.5 at T=6 / 1.0 at T=3, so per-frame
    # numbers from a short clip FAVOUR the CPU; so does the pixel extrapolation of a reduced frame
    fps = t_s / el * (hs * ws) / (h * w)
    geo = f"{hs}x{ws}" + ("" if (hs, ws) == (h, w) else f" (extrapolated by LR pixels to {h}x{w})")
    what = (f"full x4 forward of a {t_s}-frame {geo} clip, {len(secs)} x {el:.1f} s on {torch.get_num_threads()} "
            f"threads (fp32, ATen grid_sample + torchvision CPU DCNv2; frames/s of the short clip: the 30-frame clip "
            f"of the metric costs more per frame)")
    return fps, what, secs, (t_s, hs, ws)


def time_train_sample(t_s=5, size=64, state_dict=None):
    """CPU baseline of the training step on a BOUNDED sample: forward + L1 + backward (autograd through ATen
    grid_sample and torchvision's CPU DCNv2, what `--gpu_ids -1` training would execute) of ONE `t_s`-frame
    64x64 crop, fp32, all host threads.  Returns (LR frames/s, description)."""
    from eavsr_b200.synthetic import clip_inputs
    sd = state_dict or _seeded_state_dict()
    leaves = [v.requires_grad_() for k, v in sd.items() if v.is_floating_point() and not k.startswith("spynet.")
              and not k.endswith(("mean", "std", "regular_matrix"))]
    lrs = clip_inputs(1, t_s, size, size, seed=77)
    hr = F.interpolate(lrs[0], scale_factor=4, mode="bicubic", align_corners=False).clamp(0, 1).unsqueeze(0)
    eavsrp_forward(sd, clip_inputs(1, 3, size, size, seed=78), 4, "aten").mean().backward()         # lazy inits
    t0 = time.time()
    loss = (eavsrp_forward(sd, lrs, 4, "aten") - hr).abs().mean()
    loss.backward()
    el = time.time() - t0
    assert all(v.grad is not None for v in leaves[:4])
    return t_s / el, (f"forward + L1 + backward of one {t_s}-frame {size}x{size} crop in {el:.1f} s on "
                      f"{torch.get_num_threads()} threads (fp32 autograd, ATen grid_sample + torchvision CPU DCNv2); the "
                      f"15-frame crops of the metric cost more per frame")
