"""Seeded synthetic inputs shared by the CPU and GPU tests."""
import torch


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def dcn_inputs(n, cin, h, w, cout, dg, k=3, seed=0, sigma=2.0, groups=1, ho=None, wo=None):
    g = gen(seed)
    ho = h if ho is None else ho
    wo = w if wo is None else wo
    K = k * k
    x = torch.randn(n, cin, h, w, generator=g)
    off = (torch.randn(n, dg * 2 * K, ho, wo, generator=g) * sigma).clamp(-12, 12)
    r = torch.rand(off.shape, generator=g)
    off = torch.where(r < 0.02, torch.full_like(off, 300.0) * torch.sign(off), off)      # far out of bounds
    off = torch.where((r >= 0.02) & (r < 0.04), off.round(), off)                        # exact integers
    mask = torch.sigmoid(torch.randn(n, dg * K, ho, wo, generator=g))
    bound = 1.0 / (cin // groups * K) ** 0.5
    wgt = (torch.rand(cout, cin // groups, k, k, generator=g) * 2 - 1) * bound
    bias = torch.randn(cout, generator=g) * 0.1
    return x, off, mask, wgt, bias


def warp_inputs(n, c, h, w, seed=0, sigma=3.0, layout="n2hw"):
    g = gen(seed)
    x = torch.randn(n, c, h, w, generator=g)
    flow = torch.randn(n, 2, h, w, generator=g) * sigma
    r = torch.rand(flow.shape, generator=g)
    flow = torch.where(r < 0.02, flow * 200, flow)       # wildly out of range
    flow = torch.where((r >= 0.02) & (r < 0.04), flow.round(), flow)
    if layout == "nhw2":
        flow = flow.permute(0, 2, 3, 1).contiguous()
    return x, flow


def rel_err(a, ref):
    a = a.double().cpu()
    ref = ref.double().cpu()
    return ((a - ref).abs().max() / ref.pow(2).mean().sqrt().clamp_min(1e-12)).item()


def max_err(a, ref):
    return (a.double().cpu() - ref.double().cpu()).abs().max().item()


from eavsr_b200.synthetic import clip_inputs, seeded_parameters  # noqa: E402,F401


def smooth_flow_mask(flow, layout="n2hw", eps=1e-3):
    """1 where the sampling point is not within eps of an integer coordinate.  The bilinear
    kernel is only piecewise differentiable: exactly on a grid line the flow gradient is one-sided
    and the reference's fp32 normalise/un-normalise round trip lands on either side at random,
    so flow gradients are compared away from grid lines only.  Shape of ``flow``."""
    f = flow.detach().double().cpu()
    fr = f - torch.floor(f)          # grid positions are integers, so frac(x+flow) == frac(flow)
    ok = (fr > eps) & (fr < 1 - eps)
    both = ok.all(dim=1, keepdim=True) if layout == "n2hw" else ok.all(dim=3, keepdim=True)
    return both.expand_as(f).to(torch.float64)


def dcn_offset_grad_mask(off, k=3, pad=1):
    """0 for samples that sit exactly on p == -1.  There the forward value is 0 for everybody, but
    the one-sided offset gradient differs between DCNv2 implementations: mmcv's
    dmcn_get_coordinate_weight returns 0 (sample rejected, the semantics the oracle and the CUDA
    kernels follow), torchvision -- the stand-in that produced the fixtures -- returns the right
    derivative.  A measure-zero set that only the 'exact integer offsets' in the test data hit."""
    n, ch, h, w = off.shape
    o = off.detach().double().cpu().reshape(n, -1, k * k, 2, h, w)
    ii = torch.arange(k * k) // k
    jj = torch.arange(k * k) % k
    py = torch.arange(h).view(1, 1, 1, h, 1) - pad + ii.view(1, 1, -1, 1, 1) + o[:, :, :, 0]
    px = torch.arange(w).view(1, 1, 1, 1, w) - pad + jj.view(1, 1, -1, 1, 1) + o[:, :, :, 1]
    ok = ((py != -1) & (px != -1)).unsqueeze(3).expand_as(o)
    return ok.reshape(n, ch, h, w).to(torch.float64)
