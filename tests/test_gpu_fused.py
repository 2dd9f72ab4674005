"""GPU parity of the fused producer/consumer kernels against the PyTorch modules they replace
(the reference's own definitions: nn.Conv2d groups / LeakyReLU, matmul expansion, CALayer)."""
import pytest
import torch
import torch.nn.functional as F

import eavsr_b200 as E

import eavsr_b200.model as M
from eavsr_b200 import ops
from oracle import alignment as O

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("shape", [(1, 17, 23), (2, 8, 16), (1, 67, 120), (1, 5, 3)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 3e-2)])
def test_adapt_mix_matches_grouped_convs(cuda, shape, dtype, tol):
    n, h, w = shape
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn(n, 64, h, w, generator=g), torch.randn(n, 64, h, w, generator=g)
    w1, b1 = torch.randn(128, 1, 3, 3, generator=g) * 0.3, torch.randn(128, generator=g) * 0.1
    w2, b2 = torch.randn(64, 2, 3, 3, generator=g) * 0.3, torch.randn(64, generator=g) * 0.1
    cast = lambda t: t.to(dtype).double()        # noqa: E731  reference in fp64 on the dtype-rounded inputs
    y = F.leaky_relu(F.conv2d(torch.cat([cast(a), cast(b)], 1), cast(w1), cast(b1), padding=1, groups=128), 0.2)
    ref = F.leaky_relu(F.conv2d(y, cast(w2), cast(b2), padding=1, groups=64), 0.2)
    out = ops.adapt_mix(_cl(a.to(cuda, dtype)), _cl(b.to(cuda, dtype)), w1.to(cuda, dtype), b1.to(cuda, dtype),
                        w2.to(cuda, dtype), b2.to(cuda, dtype), 0.2)
    assert out.shape == ref.shape and out.dtype == dtype
    err = (out.double().cpu() - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("D", [1, 8])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cl", [True, False])
def test_affine_offsets_mask_matches_oracle(cuda, D, dtype, cl):
    g = torch.Generator().manual_seed(2)
    n, h, w = 2, 13, 19
    T = torch.randn(n, 4 * D, h, w, generator=g).to(dtype)
    t = torch.randn(n, 2 * D, h, w, generator=g).to(dtype)
    m = torch.randn(n, 9 * D, h, w, generator=g).to(dtype)
    ref_off = O.affine_offsets(T.double(), t.double(), D)
    ref_mask = torch.sigmoid(m.double())
    f = _cl if cl else (lambda z: z)
    off, mask = ops.affine_offsets_mask(f(T.to(cuda)), f(t.to(cuda)), f(m.to(cuda)), D)
    assert off.dtype == torch.float32 and off.is_contiguous() and mask.is_contiguous()
    assert (off.double().cpu() - ref_off).abs().max() < 1e-5
    assert (mask.double().cpu() - ref_mask).abs().max() < 1e-5
    off2, none = ops.affine_offsets_mask(f(T.to(cuda)), f(t.to(cuda)), None, D)
    assert none is None and torch.equal(off2, off)


@pytest.mark.parametrize("shape", [(1, 20, 30), (3, 7, 9), (1, 67, 120)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_ca_residual_matches_rcablock_tail(cuda, shape, dtype, tol):
    n, h, w = shape
    g = torch.Generator().manual_seed(3)
    res, skip = torch.randn(n, 64, h, w, generator=g), torch.randn(n, 64, h, w, generator=g)
    ca = M._CALayer(64).double()
    for p in ca.parameters():
        p.data = (torch.randn(p.shape, generator=g) * 0.5).to(dtype).double()
    ref = ca(res.to(dtype).double()) + skip.to(dtype).double()
    du = ca.conv_du
    out = ops.ca_residual(_cl(res.to(cuda, dtype)), _cl(skip.to(cuda, dtype)), du[0].weight.to(cuda, dtype),
                          du[0].bias.to(cuda, dtype), du[2].weight.to(cuda, dtype), du[2].bias.to(cuda, dtype), 16)
    assert (out.double().cpu() - ref).abs().max() < tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_ca_residual_with_folded_conv_bias(cuda, dtype, tol):
    g = torch.Generator().manual_seed(4)
    res, skip = torch.randn(2, 64, 9, 14, generator=g), torch.randn(2, 64, 9, 14, generator=g)
    rb = torch.randn(64, generator=g) * 0.5
    ca = M._CALayer(64).double()
    for p in ca.parameters():
        p.data = (torch.randn(p.shape, generator=g) * 0.5).to(dtype).double()
    r = res.to(dtype).double() + rb.to(dtype).double().view(1, -1, 1, 1)
    ref = ca(r) + skip.to(dtype).double()
    du = ca.conv_du
    out = ops.ca_residual(_cl(res.to(cuda, dtype)), _cl(skip.to(cuda, dtype)), du[0].weight.to(cuda, dtype),
                          du[0].bias.to(cuda, dtype), du[2].weight.to(cuda, dtype), du[2].bias.to(cuda, dtype), 16,
                          res_bias=rb.to(cuda, dtype))
    assert (out.double().cpu() - ref).abs().max() < tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("slope", [1.0, 0.0, 0.1])
@pytest.mark.parametrize("dtype,c", [(torch.float32, 64), (torch.bfloat16, 72), (torch.float32, 4), (torch.bfloat16, 256)])
def test_bias_act_and_conv_epilogue(cuda, slope, dtype, c):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, c, 11, 13, generator=g).to(dtype)
    b = torch.randn(c, generator=g).to(dtype)
    t = x.double() + b.double().view(1, -1, 1, 1)
    ref = torch.where(t > 0, t, t * slope)
    out = ops.bias_act_(_cl(x.to(cuda)).clone(memory_format=torch.channels_last), b.to(cuda), slope)
    tol = 1e-6 if dtype == torch.float32 else 2e-2
    assert (out.double().cpu() - ref).abs().max() < tol * max(1.0, ref.abs().max().item())
    conv = torch.nn.Conv2d(16, c, 3, 1, 1).to(cuda, dtype).to(memory_format=torch.channels_last)
    xin = _cl(torch.randn(1, 16, 10, 12, generator=g).to(cuda, dtype))
    with torch.no_grad():
        fused = ops.conv2d_bias_act(conv, xin, slope)
        y = conv(xin)
        plain = y if slope == 1.0 else F.leaky_relu(y, slope)
    assert (fused.float() - plain.float()).abs().max() < (1e-5 if dtype == torch.float32 else 5e-2)


def test_affine_offsets_with_folded_biases(cuda):
    g = torch.Generator().manual_seed(6)
    D, n, h, w = 8, 1, 9, 11
    T, t, m = (torch.randn(n, k * D, h, w, generator=g) for k in (4, 2, 9))
    bT, bt, bm = (torch.randn(k * D, generator=g) for k in (4, 2, 9))
    v = lambda z, b: z.double() + b.double().view(1, -1, 1, 1)      # noqa: E731
    ref_off = O.affine_offsets(v(T, bT), v(t, bt), D)
    ref_mask = torch.sigmoid(v(m, bm))
    off, mask = ops.affine_offsets_mask(_cl(T.to(cuda)), _cl(t.to(cuda)), _cl(m.to(cuda)), D, bT.to(cuda), bt.to(cuda),
                                        bm.to(cuda))
    assert (off.double().cpu() - ref_off).abs().max() < 1e-5
    assert (mask.double().cpu() - ref_mask).abs().max() < 1e-5


@pytest.mark.parametrize("shape", [(1, 12, 30), (2, 7, 33), (1, 67, 120), (1, 272, 480), (3, 4, 5)])
@pytest.mark.parametrize("slope", [1.0, 0.0, 0.1])
def test_conv3x3_tcgen05_matches_conv2d(cuda, shape, slope):
    n, h, w = shape
    g = torch.Generator().manual_seed(7)
    conv = torch.nn.Conv2d(64, 64, 3, 1, 1)
    with torch.no_grad():
        conv.weight.copy_((torch.rand(conv.weight.shape, generator=g) * 2 - 1) / 24)
        conv.bias.copy_(torch.randn(64, generator=g) * 0.1)
    conv = conv.to(cuda, torch.bfloat16)
    x = _cl(torch.randn(n, 64, h, w, generator=g).to(cuda, torch.bfloat16))
    with torch.no_grad():
        assert ops.conv3x3_64_eligible(conv, x)
    assert not ops.conv3x3_64_eligible(conv, x)          # gradients enabled: PyTorch path
    y = F.conv2d(x.double().cpu(), conv.weight.double().cpu(), conv.bias.double().cpu(), padding=1)
    ref = torch.where(y > 0, y, y * slope)
    with torch.no_grad():
        out, sums = ops.conv3x3_64(conv, x, slope, want_sums=True)
        out2 = ops.conv3x3_64(conv, x, slope)
    assert out.shape == ref.shape and out.dtype == torch.bfloat16 and torch.equal(out, out2)
    assert (out.double().cpu() - ref).abs().max() < 2e-2 * max(1.0, ref.abs().max().item())
    ref_sums = ref.sum((2, 3))
    assert (sums.double().cpu() - ref_sums).abs().max() < 1e-3 * max(1.0, ref_sums.abs().max().item()) + 0.02 * (h * w) ** 0.5
    with torch.no_grad():
        assert not ops.conv3x3_64_eligible(conv, x.float())


def test_rcablock_tcgen05_path_matches_torch_path(cuda):
    g = torch.Generator().manual_seed(8)
    blk = M._RCABlock(64)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * (0.05 if p.dim() > 1 else 0.1))
    blk = blk.to(cuda, torch.bfloat16).to(memory_format=torch.channels_last)
    x = _cl(torch.randn(2, 64, 21, 37, generator=g).to(cuda, torch.bfloat16))
    with torch.no_grad():
        fused = blk(x)
    ref = blk.double().cpu()(x.double().cpu())
    assert (fused.double().cpu() - ref).abs().max() < 3e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_bias_act_shuffle_matches_pixel_shuffle(cuda, dtype):
    g = torch.Generator().manual_seed(9)
    conv = torch.nn.Conv2d(64, 256, 3, 1, 1).to(cuda, dtype).to(memory_format=torch.channels_last)
    x = _cl(torch.randn(2, 64, 9, 13, generator=g).to(cuda, dtype))
    with torch.no_grad():
        fused = ops.conv2d_bias_act_shuffle(conv, x, 0.1)
        ref = F.leaky_relu(F.pixel_shuffle(conv(x), 2), 0.1)
    assert fused.shape == ref.shape == (2, 64, 18, 26)
    assert fused.is_contiguous(memory_format=torch.channels_last)
    assert (fused.float() - ref.float()).abs().max() < (1e-5 if dtype == torch.float32 else 5e-2)


def test_model_uses_fused_path_only_without_grad(cuda):
    blk = M._RCABlock(64).to(cuda)
    x = _cl(torch.randn(1, 64, 12, 12, device=cuda))
    from eavsr_b200 import _lib
    n0 = _lib.launch_count()
    with torch.no_grad():
        y_fused = blk(x)
    assert _lib.launch_count() - n0 == 3        # bias+ReLU epilogue, channel sums, scale+residual
    n1 = _lib.launch_count()
    y_torch = blk(x.requires_grad_())
    # with autograd on, none of the fused (non-differentiable) kernels runs: cuDNN convolutions + the differentiable
    # channel pooling of the attention layer (one channel-sum launch)
    assert _lib.launch_count() - n1 == 1 and y_torch.requires_grad
    n2 = _lib.launch_count()
    assert (y_fused - y_torch.detach()).abs().max() < 1e-4
    y_torch.sum().backward()
    assert x.grad is not None and all(p.grad is not None for p in blk.parameters())
    assert _lib.launch_count() - n2 == 3        # backward: two bias gradients + the attention scale's channel-dot


@pytest.mark.parametrize("shape", [(1, 40, 72), (2, 37, 53), (1, 270, 480)])
def test_dcn_affine_equals_expansion_then_dcn(cuda, shape):
    """the DCN kernel that expands T*R - R + t and sigmoid(logits) itself equals
    eavsr_affine_offsets_forward followed by eavsr_dcn_forward on the same bf16 operands (same fp32
    expansion, same gather; only the window apron differs), and stays inside the bf16 tolerance of the
    fp64 oracle of the composition."""
    from eavsr_b200 import ops
    from oracle import alignment as O
    n, h, w = shape
    D = 8
    g = torch.Generator().manual_seed(41)
    x = torch.randn(n, 64, h, w, generator=g).bfloat16()
    aff = torch.randn(n, 15 * D, h, w, generator=g)
    aff[:, :4 * D] = aff[:, :4 * D] * 0.3 + torch.tensor([1.0, 0.0, 0.0, 1.0]).repeat(D).view(1, -1, 1, 1)
    aff[:, 4 * D:6 * D] *= 1.5
    aff = aff.bfloat16()
    ab = (torch.randn(15 * D, generator=g) * 0.2).bfloat16()
    wgt = ((torch.rand(64, 64, 3, 3, generator=g) * 2 - 1) / 24).bfloat16()
    bias = (torch.randn(64, generator=g) * 0.1).bfloat16()
    xd = x.to(cuda).contiguous(memory_format=torch.channels_last)
    ad = aff.to(cuda).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        assert ops.dcn_affine_eligible(xd, ad, wgt.to(cuda), D)
        fused = ops.dcn_affine(xd, ad, ab.to(cuda), wgt.to(cuda), bias.to(cuda), D)
        off, msk = ops.affine_offsets_mask(ad[:, :4 * D], ad[:, 4 * D:6 * D], ad[:, 6 * D:], D, ab[:4 * D].to(cuda),
                                           ab[4 * D:6 * D].to(cuda), ab[6 * D:].to(cuda))
        comp = ops.modulated_deform_conv2d(xd, off, msk, wgt.to(cuda), bias.to(cuda), 1, 1, 1, 1, D)
    assert fused.shape == comp.shape and fused.dtype == torch.bfloat16
    # written in place into the middle 64 channels of a wider NHWC buffer (the caller's torch.cat input)
    wide = torch.full((n, 192, h, w), 7.0, dtype=torch.bfloat16, device=cuda).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        got = ops.dcn_affine(xd, ad, ab.to(cuda), wgt.to(cuda), bias.to(cuda), D, out=wide[:, 64:128])
    assert got.data_ptr() == wide[:, 64:128].data_ptr() and torch.equal(wide[:, 64:128], fused)
    assert bool((wide[:, :64] == 7).all()) and bool((wide[:, 128:] == 7).all())
    rms = comp.float().pow(2).mean().sqrt().item()
    assert (fused.float() - comp.float()).abs().max().item() < 2e-2 * rms
    assert (fused.float() - comp.float()).abs().mean().item() < 2e-4 * rms     # identical up to rare rounding flips
    T = aff[:, :4 * D].double() + ab[:4 * D].double().view(1, -1, 1, 1)
    t = aff[:, 4 * D:6 * D].double() + ab[4 * D:6 * D].double().view(1, -1, 1, 1)
    lg = aff[:, 6 * D:].double() + ab[6 * D:].double().view(1, -1, 1, 1)
    ref = O.modulated_deform_conv2d(x.double(), O.affine_offsets(T, t, D), torch.sigmoid(lg), wgt.double(),
                                    bias.double(), 1, 1, 1, 1, D)
    assert ((fused.double().cpu() - ref).abs().max() / ref.pow(2).mean().sqrt()).item() < 3e-2


@pytest.mark.parametrize("shape", [(1, 24, 60), (2, 37, 53), (1, 272, 480)])
def test_conv3x3_with_fused_channel_attention(cuda, shape):
    """eavsr_conv3x3_ca_forward == eavsr_ca_scale_forward followed by eavsr_conv3x3_forward, bit for bit
    (same MLP, same fp32 scale-and-add, same bf16 rounding of y), including the y it hands to the next block."""
    n, h, w = shape
    g = torch.Generator().manual_seed(51)
    conv0 = torch.nn.Conv2d(64, 64, 3, 1, 1).to(cuda, torch.bfloat16)
    conv1 = torch.nn.Conv2d(64, 64, 3, 1, 1).to(cuda, torch.bfloat16)
    du = M._CALayer(64).to(cuda, torch.bfloat16).conv_du
    skip = torch.randn(n, 64, h, w, generator=g).to(cuda, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    hin = torch.randn(n, 64, h, w, generator=g).to(cuda, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        res, sums = ops.conv3x3_64(conv0, hin, 1.0, want_sums=True)
        y_ref = ops.ca_scale(res, skip, sums, du[0].weight, du[0].bias, du[2].weight, du[2].bias, 16)
        out_ref, s_ref = ops.conv3x3_64(conv1, y_ref, 0.0, want_sums=True)
        out, y, s2 = ops.conv3x3_64_ca(conv1, skip, res, sums, du[0].weight, du[0].bias, du[2].weight, du[2].bias,
                                       0.0, want_sums=True)
    assert torch.equal(y, y_ref)
    assert torch.equal(out, out_ref)
    assert torch.allclose(s2, s_ref, rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("pad", ["zeros", "border"])
def test_flow_warp2_equals_two_warps(cuda, pad):
    """Row f2: both 64-channel maps warped with one flow in one launch == two eavsr_flow_warp_forward calls."""
    g = torch.Generator().manual_seed(61)
    mk = lambda: torch.randn(2, 64, 37, 53, generator=g).to(cuda, torch.bfloat16).contiguous(memory_format=torch.channels_last)   # noqa: E731
    a, b = mk(), mk()
    flow = (torch.randn(2, 2, 37, 53, generator=g) * 3).to(cuda)
    with torch.no_grad():
        o1, o2 = ops.flow_warp2(a, b, flow, padding_mode=pad)
        r1, r2 = ops.flow_warp(a, flow, padding_mode=pad), ops.flow_warp(b, flow, padding_mode=pad)
    assert torch.equal(o1, r1) and torch.equal(o2, r2)
    # not eligible (fp32): falls back to two calls
    o1, o2 = ops.flow_warp2(a.float(), b.float(), flow)
    assert torch.allclose(o1, ops.flow_warp(a.float(), flow)) and o2.dtype == torch.float32


@pytest.mark.parametrize("nb,shape", [(1, (1, 24, 60)), (3, (2, 37, 53)), (5, (1, 68, 120)), (30, (1, 272, 480))])
def test_rca_group_chain_equals_launch_per_convolution(cuda, nb, shape):
    """Row f3: the whole RCAGroup in ONE cooperative launch (eavsr_conv3x3_chain_forward: grid barrier between
    layers, recycled activation buffers read through L2) == the launch-per-convolution tcgen05 path, and both ==
    the PyTorch modules in fp64 up to bf16 rounding."""
    from eavsr_b200 import _lib
    n, h, w = shape
    g = torch.Generator().manual_seed(70 + nb)
    grp = M._RCAGroup(64, nb)
    with torch.no_grad():
        for p in grp.parameters():
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * (0.04 if p.dim() > 1 else 0.1))
    ref_mod = grp.double() if nb <= 5 else None
    x = torch.randn(n, 64, h, w, generator=g)
    ref = None
    if ref_mod is not None:
        with torch.no_grad():
            y = x.bfloat16().double()
            for blk in ref_mod.rg[:-1]:
                y = blk.ca(blk.res(y)) + y
            ref = ref_mod.rg[-1](y) + x.bfloat16().double()
    grp = grp.to(cuda, torch.bfloat16).to(memory_format=torch.channels_last)
    xb = _cl(x.to(cuda, torch.bfloat16))
    with torch.no_grad():
        grp.chain = False
        grp(xb)                              # first call packs the weights (one extra launch per convolution)
        n0 = _lib.launch_count()
        per_conv = grp(xb)
        n1 = _lib.launch_count()
        grp.chain = True
        chained = grp(xb)
        n2 = _lib.launch_count()
        again = grp(xb)                      # recycled buffers / re-armed grid barrier: a second call is the same
    assert n1 - n0 == 2 * nb + 1 and n2 - n1 == 1
    rms = per_conv.float().pow(2).mean().sqrt().item()
    # the channel sums are accumulated with atomics in a different order (more so with two epilogue teams): equal up
    # to that rounding -- worst element within 2 % of the RMS or two bf16 ulps of its own magnitude, bulk within 0.2 %
    def close(a, b):
        a, b = a.float(), b.float()
        return bool(((a - b).abs() <= torch.maximum(torch.full_like(b, 2e-2 * rms), b.abs() * 2.0 ** -6)).all())
    assert close(chained, per_conv)
    assert (chained.float() - per_conv.float()).pow(2).mean().sqrt().item() <= 2e-3 * rms
    assert close(again, chained)
    assert (again.float() - chained.float()).pow(2).mean().sqrt().item() <= 2e-3 * rms
    if ref is not None:
        assert (chained.double().cpu() - ref).abs().max().item() < 3e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("hw", [(68, 120), (34, 60), (16, 24), (272, 480)])
def test_flow_warp_pyramid_equals_interpolate_add_warp(cuda, dtype, hw):
    """Row f2: the interpolate -> scale -> add -> warp chains of MultiAdSTN.forward (models/networks.py:600-615,619)
    evaluated inside the warp kernel == F.interpolate(align_corners=True) + elementwise + flow_warp."""
    h, w = hw
    g = torch.Generator().manual_seed(90)
    up = lambda t, size: F.interpolate(t, size=size, mode="bilinear", align_corners=True)     # noqa: E731
    x = _cl(torch.randn(2, 64, h, w, generator=g).to(cuda, dtype))
    x2 = _cl(torch.randn(2, 64, h, w, generator=g).to(cuda, dtype))
    full = (torch.randn(2, 2, 4 * h, 4 * w, generator=g) * 6).to(cuda)           # a flow 4x finer (level 3 of the pyramid)
    same = (torch.randn(2, 2, h, w, generator=g) * 2).to(cuda)
    half = (torch.randn(2, 2, h // 2, w // 2, generator=g) * 1.5).to(cuda)        # coarser fields (levels 2 / 1)
    half2 = (torch.randn(2, 2, h // 2, w // 2, generator=g) * 1.5).to(cuda)
    with torch.no_grad():
        assert ops.flow_warp_pyramid_eligible(x)
        # level 3: one 4x finer flow, scaled by 1/4
        ref_flow = up(full, (h, w)) / 4.
        out = ops.flow_warp_pyramid(x, [(full, 0.25)])
        assert torch.equal(out, E.flow_warp(x, ref_flow)) or (out.float() - E.flow_warp(x, ref_flow).float()).abs().max() < 2e-2
        # level 2 / 1: several terms, one kept, the sum returned
        t1, t2 = up(half, (h, w)) * 2, up(half2, (h, w)) * 2
        ref_flow = same + t1 + t2
        out, fl, kept = ops.flow_warp_pyramid(x, [(same, 1.0), (half, 2.0), (half2, 2.0)], want_flow=True, keep=(2,))
        assert (fl - ref_flow).abs().max().item() < 1e-4 and (kept - t2).abs().max().item() < 1e-5
        ref = E.flow_warp(x, fl)                       # same flow tensor -> the warp itself must agree exactly
        assert torch.equal(out, ref)
        if dtype == torch.bfloat16:                    # dual warp with a 2-term flow (the final warp of MultiAdSTN)
            o1, o2 = ops.flow_warp_pyramid(x, [(fl, 1.0), (same, 1.0)], x2=x2)
            r1, r2 = ops.flow_warp2(x, x2, fl + same)
            assert torch.equal(o1, r1) and torch.equal(o2, r2)


def test_spynet_level_input_equals_interpolate_warp_cat(cuda):
    """Row f4: one launch builds cat[ref, flow_warp(supp, up, 'border'), up] with up = 2 * resize_x2(flow_prev)
    (SPyNet.compute_flow, models/eavsrp_model.py:468-486)."""
    g = torch.Generator().manual_seed(91)
    for (n, h, w) in ((3, 18, 30), (2, 9, 15), (29, 72, 120)):
        ref = torch.rand(n, 3, h, w, generator=g).to(cuda)
        supp = torch.rand(n, 3, h, w, generator=g).to(cuda)
        with torch.no_grad():
            x0 = ops.spynet_level_input(ref, supp, None)
            want0 = torch.cat([ref, supp, torch.zeros(n, 2, h, w, device=cuda)], 1)
            assert torch.equal(x0, want0)                                  # zero flow: the border warp is the identity
            if h % 2 == 0:
                prev = (torch.randn(n, 2, h // 2, w // 2, generator=g) * 2).to(cuda)
                up = F.interpolate(prev, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
                want = torch.cat([ref, E.flow_warp_nhw2(supp, up.permute(0, 2, 3, 1), padding_mode="border"), up], 1)
                got = ops.spynet_level_input(ref, supp, prev)
                assert (got[:, 6:] - up).abs().max().item() < 1e-5
                assert (got - want).abs().max().item() < 1e-4


def test_multi_adstn_pyramid_folding_matches_unfolded_path(cuda):
    g = torch.Generator().manual_seed(92)
    m = M.MultiAdSTN(64, 8)
    from eavsr_b200.synthetic import seeded_parameters
    seeded_parameters(m)
    m = m.to(cuda, torch.bfloat16).to(memory_format=torch.channels_last)
    h, w = 48, 80
    feats = lambda: [_cl(torch.randn(1, 64, h >> i, w >> i, generator=g).to(cuda, torch.bfloat16)) for i in range(3)]   # noqa: E731
    nbr, ref, prop = feats(), feats(), feats()[0]
    flow = (torch.randn(1, 2, h, w, generator=g) * 2).to(cuda)
    with torch.no_grad():
        m.fold_pyramid = True
        a = m(nbr, ref, prop, flow).float()
        m.fold_pyramid = False
        b = m(nbr, ref, prop, flow).float()
    rms = b.pow(2).mean().sqrt().item()
    assert (a - b).pow(2).mean().sqrt().item() < 1e-2 * rms        # same arithmetic up to fp32 summation order of the flows


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_cat_channels_equals_torch_cat(cuda, dtype):
    """nhwc_cat (the concatenations in front of the fusion / backbone / reconstruction convolutions,
    models/eavsrp_model.py:271-324, 350-364): bit-exact copy, whole buffer and channel slice, odd pixel counts,
    and the torch.cat fallback when a gradient is needed or the layout is not channels_last."""
    from eavsr_b200 import ops
    g = torch.Generator().manual_seed(5)
    for n, h, w, cs in ((2, 19, 23, (64, 64, 128)), (1, 270, 480, (64, 64, 64, 64, 64)), (1, 5, 7, (8,))):
        ts = [_cl(torch.randn(n, c, h, w, generator=g).to(cuda, dtype)) for c in cs]
        with torch.no_grad():
            out = ops.cat_channels(ts)
            assert out.is_contiguous(memory_format=torch.channels_last) and torch.equal(out, torch.cat(ts, 1))
            buf = torch.full((n, sum(cs) + 16, h, w), 7.0, device=cuda, dtype=dtype).contiguous(memory_format=torch.channels_last)
            ops.cat_channels(ts, out=buf, channel_offset=8)
            assert torch.equal(buf[:, 8:8 + sum(cs)], torch.cat(ts, 1))
            assert bool((buf[:, :8] == 7).all()) and bool((buf[:, 8 + sum(cs):] == 7).all())
    a = torch.randn(1, 8, 6, 6, device=cuda, requires_grad=True)
    b = torch.randn(1, 8, 6, 6, device=cuda)
    y = ops.cat_channels([a, b])                                       # autograd on: torch.cat
    y.sum().backward()
    assert a.grad is not None and torch.equal(y, torch.cat([a, b], 1))
    with torch.no_grad():                                              # NCHW-contiguous inputs: torch.cat
        c = torch.randn(2, 12, 6, 6, device=cuda)
        assert torch.equal(ops.cat_channels([c, c]), torch.cat([c, c], 1))


@pytest.mark.parametrize("cin,cout", [(128, 128), (128, 64), (16, 16), (32, 16)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_grouped_conv3x3_forward_backward_match_conv2d(cuda, cin, cout, dtype, tol):
    """Row a3, training path: the depthwise / two-inputs-per-group 3x3 convolutions of the AdaptBlocks
    (models/networks.py:289-290, 327-328) on the library's kernels == nn.Conv2d in fp64: output, d(input),
    d(weight), d(bias); odd sizes, batch > 1, with and without bias, and under torch.autocast."""
    from eavsr_b200 import ops
    g = torch.Generator().manual_seed(cin + cout)
    for n, h, w, bias in ((2, 13, 19, True), (1, 64, 64, True), (3, 5, 8, False)):
        conv = torch.nn.Conv2d(cin, cout, 3, 1, 1, groups=cout, bias=bias)
        with torch.no_grad():
            for p in conv.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
        x = torch.randn(n, cin, h, w, generator=g)
        go = torch.randn(n, cout, h, w, generator=g)
        ref_conv = torch.nn.Conv2d(cin, cout, 3, 1, 1, groups=cout, bias=bias).double()
        ref_conv.load_state_dict(conv.state_dict())
        if dtype == torch.bfloat16:                            # reference on the bf16-rounded operands
            with torch.no_grad():
                for p in ref_conv.parameters():
                    p.copy_(p.float().bfloat16().double())
        xr = x.to(dtype).double().requires_grad_()
        yr = ref_conv(xr)
        yr.backward(go.to(dtype).double())
        conv = conv.to(cuda, dtype)
        assert ops.grouped_conv3x3_eligible(conv, x.to(cuda))
        xg = _cl(x.to(cuda, dtype)).requires_grad_()
        y = ops.grouped_conv3x3(conv, xg)
        assert y.dtype == dtype and y.is_contiguous(memory_format=torch.channels_last)
        y.backward(_cl(go.to(cuda, dtype)))
        scale = lambda t: max(1.0, t.abs().max().item())      # noqa: E731
        assert (y.double().cpu() - yr).abs().max().item() <= tol * scale(yr)
        assert (xg.grad.double().cpu() - xr.grad).abs().max().item() <= tol * scale(xr.grad)
        assert (conv.weight.grad.double().cpu() - ref_conv.weight.grad).abs().max().item() <= tol * scale(ref_conv.weight.grad)
        if bias:
            assert (conv.bias.grad.double().cpu() - ref_conv.bias.grad).abs().max().item() <= tol * scale(ref_conv.bias.grad)
    conv = torch.nn.Conv2d(cin, cout, 3, 1, 1, groups=cout).to(cuda)          # fp32 parameters, bf16 autocast
    x = torch.randn(1, cin, 9, 11, device=cuda, requires_grad=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = ops.grouped_conv3x3(conv, x)
        yt = conv(x)
    assert y.dtype == torch.bfloat16 and (y.float() - yt.float()).abs().max().item() <= 3e-2 * max(1.0, yt.abs().max().item())
    y.float().sum().backward()
    assert conv.weight.grad.dtype == torch.float32 and x.grad.dtype == torch.float32


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_native_bias_gradient_and_channel_mean_match_torch(cuda, dtype, tol):
    """Training path of the residual backbone: `conv2d_native_bias_grad` (cuDNN convolution, bias gradient on the
    channel-sum kernel) and `channel_mean` (CALayer's global pooling, models/networks.py:431-447) == the PyTorch ops."""
    from eavsr_b200 import ops
    g = torch.Generator().manual_seed(9)
    conv = torch.nn.Conv2d(64, 64, 3, 1, 1).to(cuda, dtype).to(memory_format=torch.channels_last)
    ref = torch.nn.Conv2d(64, 64, 3, 1, 1).to(cuda).double()
    ref.load_state_dict({k: v.double() for k, v in conv.state_dict().items()})
    for n, h, w in ((2, 19, 23), (8, 64, 64)):
        x = _cl(torch.randn(n, 64, h, w, generator=g).to(cuda, dtype))
        go = _cl(torch.randn(n, 64, h, w, generator=g).to(cuda, dtype))
        xa, xb = x.clone().requires_grad_(), x.double().requires_grad_()
        conv.zero_grad(); ref.zero_grad()
        y = ops.conv2d_native_bias_grad(conv, xa)
        y.backward(go)
        yr = ref(xb)
        yr.backward(go.double())
        sc = lambda t: max(1.0, t.abs().max().item())         # noqa: E731
        assert (y.double() - yr).abs().max().item() <= max(tol, 2e-2 if dtype == torch.bfloat16 else tol) * sc(yr)
        assert (conv.bias.grad.double() - ref.bias.grad).abs().max().item() <= tol * sc(ref.bias.grad)
        assert (xa.grad.double() - xb.grad).abs().max().item() <= max(tol, 2e-2 if dtype == torch.bfloat16 else tol) * sc(xb.grad)
        xm, xr = x.clone().requires_grad_(), x.double().requires_grad_()
        m = ops.channel_mean(xm)
        mr = xr.mean((2, 3), keepdim=True)
        assert m.shape == mr.shape and (m.double() - mr).abs().max().item() <= tol * sc(mr)
        gm = torch.randn(m.shape, generator=g).to(cuda, dtype)
        m.backward(gm)
        mr.backward(gm.double())
        assert (xm.grad.double() - xr.grad).abs().max().item() <= tol * sc(xr.grad)
        # res * scale + skip with the native scale gradient (RCABlock tail, models/networks.py:449-465)
        r1, s1, k1 = (t.clone().requires_grad_() for t in (x, torch.rand(n, 64, 1, 1, generator=g).to(cuda, dtype), go))
        r2, s2, k2 = (t.detach().double().requires_grad_() for t in (r1, s1, k1))
        o1 = ops.scale_residual(r1, s1, k1)
        o2 = r2 * s2 + k2
        gg = _cl(torch.randn(n, 64, h, w, generator=g).to(cuda, dtype))
        o1.backward(gg)
        o2.backward(gg.double())
        btol = max(tol, 2e-2 if dtype == torch.bfloat16 else tol)
        assert (o1.double() - o2).abs().max().item() <= btol * sc(o2)
        assert (r1.grad.double() - r2.grad).abs().max().item() <= btol * sc(r2.grad)
        assert (s1.grad.double() - s2.grad).abs().max().item() <= btol * sc(s2.grad)
        assert torch.equal(k1.grad, gg)
