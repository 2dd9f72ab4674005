"""Generate the golden fixtures in tests/golden/ from the REFERENCE itself.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What is executed is the reference's own code, imported unmodified from /root/reference through
the import shims in tests/golden/_shims (mmcv / cupy are not installable here):
  * flow_warp            models/networks.py:699-739, models/eavsrp_model.py:587-626
  * BaseModel.backwarp   models/base_model.py:321-354
  * DCNv2 call site      models/networks.py:627-630 (mmcv.ops shim -> torchvision.ops.deform_conv2d)
  * AdaptBlockOffset     models/networks.py:280-315
  * MultiAdSTN.forward   models/networks.py:597-631
  * correlation kernels  pwc/correlation/correlation.py:8-233, executed under cuda_emu.py with the
                         reference's own launch geometry (:293-322, :343-373) and its own
                         cupy_kernel() size substitution (:235-271)
  * EAVSRP x4 / x2       models/eavsrp_model.py:121-364, models/eavsrpx2_model.py
Weights are not stored: they are re-created from the parameter names by helpers.seeded_parameters.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("EAVSR_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "_shims"), REF, os.path.join(ROOT, "tests"), ROOT]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torchvision.models.vgg as _vgg  # noqa: E402

_orig_vgg16 = _vgg.vgg16
_vgg.vgg16 = lambda pretrained=False, **kw: _orig_vgg16(weights=None, **kw)   # no network for ImageNet weights

from helpers import clip_inputs, dcn_inputs, seeded_parameters, warp_inputs  # noqa: E402

torch.set_grad_enabled(True)


def save(name, **arrays):
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB")


def gen_flow_warp():
    import models.networks as N
    import models.eavsrp_model as M
    for tag, fn, layout, pad in (("networks_zeros", N.flow_warp, "n2hw", "zeros"),
                                 ("model_zeros", M.flow_warp, "nhw2", "zeros"),
                                 ("model_border", M.flow_warp, "nhw2", "border")):
        x, flow = warp_inputs(2, 6, 11, 14, seed=100, layout=layout)
        x.requires_grad_()
        flow.requires_grad_()
        out = fn(x, flow, padding_mode=pad)
        g = torch.randn(out.shape, generator=torch.Generator().manual_seed(101))
        gx, gf = torch.autograd.grad(out, [x, flow], g)
        save("flow_warp_" + tag, x=x, flow=flow, out=out, g=g, gx=gx, gflow=gf)


def gen_flow_warp_size1():
    """size-1 dimensions: the reference's 2/max(size-1,1) normalisation makes the flow irrelevant there."""
    import models.networks as N
    import models.eavsrp_model as M
    out = {}
    for tag, fn, layout, shape in (("n_row", N.flow_warp, "n2hw", (1, 4, 1, 5)), ("n_col", N.flow_warp, "n2hw", (1, 4, 6, 1)),
                                   ("m_px", M.flow_warp, "nhw2", (2, 3, 1, 1))):
        x, flow = warp_inputs(*shape, seed=108, sigma=0.7, layout=layout)
        out[tag + "_x"], out[tag + "_flow"], out[tag + "_out"] = x, flow, fn(x, flow)
    save("flow_warp_size1", **out)


def gen_backwarp():
    from models.base_model import BaseModel
    dummy = types.SimpleNamespace(backwarp_tenGrid={}, backwarp_tenPartial={})
    dummy.backwarp = types.MethodType(BaseModel.backwarp, dummy)
    x, flow = warp_inputs(2, 3, 12, 16, seed=102)
    out, mask = BaseModel.get_backwarp(dummy, None, x, None, flow=flow)
    save("backwarp", x=x, flow=flow, out=out, mask=mask)


def gen_dcn():
    import models.networks as N
    for dg in (8, 16):
        x, off, mask, w, b = dcn_inputs(1, 64, 9, 11, 64, dg, seed=103)
        leaves = [t.requires_grad_() for t in (x, off, mask, w, b)]
        out = N.modulated_deform_conv2d(*leaves, (1, 1), (1, 1), (1, 1), 1, dg)
        g = torch.randn(out.shape, generator=torch.Generator().manual_seed(104))
        grads = torch.autograd.grad(out, leaves, g)
        save(f"dcn_dg{dg}", out=out, g=g, gx=grads[0], goffset=grads[1], gmask=grads[2], gweight=grads[3],
             gbias=grads[4])


def gen_adastn():
    import models.networks as N
    opt = types.SimpleNamespace(n_frame=5)
    torch.manual_seed(0)
    m = N.AdaptBlockOffset(opt, deformable_groups=8)
    seeded_parameters(m)
    g = torch.Generator().manual_seed(105)
    a = torch.randn(1, 64, 9, 12, generator=g)
    b = torch.randn(1, 64, 9, 12, generator=g)
    with torch.no_grad():
        off, mask = m(a, b)
        x_h = m.concat2(m.concat(torch.cat([a, b], 1)))
        T = m.transform_matrix_conv(x_h)
        t = m.translation_conv(x_h)
    save("adapt_block_offset", a=a, b=b, offset=off, mask=mask, transform=T, translation=t)

    m2 = N.MultiAdSTN(opt, 64, 64, deformable_groups=8)
    seeded_parameters(m2)
    h, w = 16, 24
    nbr = [torch.randn(1, 64, h >> i, w >> i, generator=g) for i in range(3)]
    ref = [torch.randn(1, 64, h >> i, w >> i, generator=g) for i in range(3)]
    prop = torch.randn(1, 64, h, w, generator=g)
    flow = torch.randn(1, 2, h, w, generator=g) * 2
    with torch.no_grad():
        out = m2(nbr, ref, prop, flow)
    save("multi_adstn", out=out, flow=flow, prop=prop, **{f"nbr{i}": nbr[i] for i in range(3)},
         **{f"ref{i}": ref[i] for i in range(3)})


def gen_correlation():
    import cuda_emu
    from pwc.correlation import correlation as C

    def run(name, variables, grid, block, args, shared=0):
        cuda_emu.launch(C.cupy_kernel(name, variables), name, grid, block, args, shared)

    for tag, shape in (("a", (2, 8, 6, 7)), ("b", (1, 35, 9, 5))):
        g = torch.Generator().manual_seed(106)
        first = torch.randn(shape, generator=g)
        second = torch.randn(shape, generator=g)
        n_, c, h, w = shape
        # forward: launch sequence of _FunctionCorrelation.forward (correlation.py:280-322)
        rbot0 = first.new_zeros([n_, h + 8, w + 8, c])
        rbot1 = first.new_zeros([n_, h + 8, w + 8, c])
        out = first.new_zeros([n_, 81, h, w])
        n = h * w
        run('kernel_Correlation_rearrange', {'input': first, 'output': rbot0},
            [int((n + 16 - 1) / 16), c, n_], [16, 1, 1], [n, first, rbot0])
        run('kernel_Correlation_rearrange', {'input': second, 'output': rbot1},
            [int((n + 16 - 1) / 16), c, n_], [16, 1, 1], [n, second, rbot1])
        n = out.shape[1] * out.shape[2] * out.shape[3]
        run('kernel_Correlation_updateOutput', {'rbot0': rbot0, 'rbot1': rbot1, 'top': out},
            [w, h, n_], [32, 1, 1], [n, rbot0, rbot1, out], shared=c * 4)
        # backward: launch sequence of _FunctionCorrelation.backward (correlation.py:332-373)
        gout = torch.randn(out.shape, generator=g)
        g1 = first.new_zeros(shape)
        g2 = first.new_zeros(shape)
        for s in range(n_):
            n = c * h * w
            run('kernel_Correlation_updateGradFirst',
                {'rbot0': rbot0, 'rbot1': rbot1, 'gradOutput': gout, 'gradFirst': g1, 'gradSecond': None},
                [int((n + 512 - 1) / 512), 1, 1], [512, 1, 1], [n, s, rbot0, rbot1, gout, g1, None])
            run('kernel_Correlation_updateGradSecond',
                {'rbot0': rbot0, 'rbot1': rbot1, 'gradOutput': gout, 'gradFirst': None, 'gradSecond': g2},
                [int((n + 512 - 1) / 512), 1, 1], [512, 1, 1], [n, s, rbot0, rbot1, gout, None, g2])
        save("correlation_" + tag, first=first, second=second, out=out, gout=gout, gfirst=g1, gsecond=g2)


def gen_pwc():
    """The reference's PWCNET (models/pwc_net.py) and BaseModel.estimate / get_backwarp (models/base_model.py:
    294-360) on CPU: the cupy cost volume cannot run here, so `correlation.FunctionCorrelation` is replaced by
    the oracle's restatement (itself pinned to the reference's kernel strings by gen_correlation); everything
    else -- Extractor / Decoder / Refiner, Decoder.backwarp, estimate, get_backwarp -- is the reference's code."""
    import models.pwc_net as P
    from models.base_model import BaseModel
    from oracle import alignment as O
    P.correlation.FunctionCorrelation = lambda tenFirst, tenSecond: O.correlation(tenFirst, tenSecond)
    # the constructor's last statement loads ./pwc/pwc-default (models/pwc_net.py:249-251), a stripped blob
    orig_load, orig_lsd = torch.load, P.PWCNET.load_state_dict
    torch.load = lambda *a, **k: {}
    P.PWCNET.load_state_dict = lambda self, *a, **k: None
    try:
        net = P.PWCNET()
    finally:
        torch.load, P.PWCNET.load_state_dict = orig_load, orig_lsd
    net.eval()
    seeded_parameters(net)
    a = clip_inputs(1, 2, 64, 64, seed=110)[0]                      # two related frames (1, 3, 64, 64) each
    first, second = a[0:1], a[1:2]
    with torch.no_grad():
        flow_net = net(first, second)
    dummy = types.SimpleNamespace(backwarp_tenGrid={}, backwarp_tenPartial={})
    for name in ("backwarp", "estimate", "get_flow", "get_backwarp"):
        setattr(dummy, name, types.MethodType(getattr(BaseModel, name), dummy))
    lr = clip_inputs(1, 1, 40, 56, seed=111)[0]                     # not a multiple of 64: estimate() resizes
    hr = F_interp(clip_inputs(1, 1, 40, 56, seed=112)[0], 2)
    with torch.no_grad():
        flow_est = dummy.estimate(lr, F_interp(hr, 0.5, True), net)
        out, mask = dummy.get_backwarp(lr, hr, net, scale=2)
    keys = sorted(net.state_dict().keys())
    save("pwc_net", first=first, second=second, flow_net=flow_net, lr=lr, hr=hr, flow_est=flow_est, out=out, mask=mask,
         keys=np.array(keys), key_shapes=np.array([str(tuple(net.state_dict()[k].shape)) for k in keys]))
    print("  pwc flow", float(flow_net.abs().mean()), "est", float(flow_est.abs().mean()), "mask", float(mask.mean()))


def F_interp(x, s, ac=False):
    import torch.nn.functional as F
    return F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=ac)


def gen_model():
    import importlib
    for scale, modname, t in ((4, "models.eavsrp_model", 4), (2, "models.eavsrpx2_model", 3)):
        M = importlib.import_module(modname)
        opt = types.SimpleNamespace(scale=scale, predict=False, n_frame=t, n_flow=5)
        torch.manual_seed(0)
        net = M.EAVSRP(opt, None).eval()
        shapes = seeded_parameters(net)
        lrs = clip_inputs(1, t, 64, 64, seed=107)
        with torch.no_grad():
            sr = net(lrs)
        keys = sorted(net.state_dict().keys())
        save(f"eavsrp_x{scale}", sr_sub=sr[..., ::4, ::4].contiguous(), sr_sub2=sr[..., 1::8, 2::8].contiguous(),
             sr_mean=sr.mean(), sr_abs_mean=sr.abs().mean(), shape=np.array(sr.shape),
             keys=np.array(keys), nparams=np.array(sum(p.numel() for p in net.parameters())),
             key_shapes=np.array([str(tuple(net.state_dict()[k].shape)) for k in keys]))
        print("  sr", tuple(sr.shape), float(sr.mean()), float(sr.abs().max()), "params", len(shapes))


if __name__ == "__main__":
    which = sys.argv[1:] or ["flow_warp", "flow_warp_size1", "backwarp", "dcn", "adastn", "correlation", "pwc", "model"]
    for w in which:
        globals()["gen_" + w]()
