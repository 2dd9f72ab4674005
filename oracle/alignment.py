"""CPU restatement (oracle) of EAVSR's inter-frame alignment operators.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Plain torch on the CPU,
written as explicit index arithmetic (no F.grid_sample, no torchvision, no
mmcv) so that it is an independent statement of the algorithm; every function
is differentiable, which gives the backward oracles through autograd.

Reference sites restated here (paths relative to /root/reference):
  * modulated_deform_conv2d  -- call site models/networks.py:627-630; algorithm
    = DCNv2 as published (Zhu et al. 2018) and as implemented by mmcv-full 1.x
    ``modulated_deformable_im2col`` + GEMM (third-party, not vendored; version
    un-pinned by README.md:24).
  * networks.flow_warp       -- models/networks.py:699-739  (flow (n,2,h,w))
  * eavsrp_model.flow_warp   -- models/eavsrp_model.py:587-626 (flow (n,h,w,2))
  * BaseModel.backwarp       -- models/base_model.py:321-354 (align_corners=False)
  * FunctionCorrelation      -- pwc/correlation/correlation.py:35-103 (forward),
                                :105-233 (backward)
  * AdaptBlockOffset offset expansion -- models/networks.py:298-315
"""
from __future__ import annotations

import torch

__all__ = [
    "bilinear_gather",
    "modulated_deform_conv2d",
    "flow_warp",
    "backwarp",
    "correlation",
    "correlation_backward",
    "affine_offsets",
]


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def bilinear_gather(x: torch.Tensor, py: torch.Tensor, px: torch.Tensor,
                    border: bool = False, dcn_reject: bool = False) -> torch.Tensor:
    """Bilinear sample of ``x (n,c,h,w)`` at pixel coordinates ``py, px (n,P)``.

    Each of the four corners contributes iff it lies inside the image (zero
    padding).  ``border`` clamps the coordinate to [0, size-1] first
    (grid_sample padding_mode='border').  ``dcn_reject`` applies DCNv2's extra
    rule that a sample with p <= -1 or p >= size is exactly zero.
    Returns (n, c, P).
    """
    n, c, h, w = x.shape
    if border:
        # grid_sampler's clip_coordinates_set_grad: the coordinate gradient is zero when the
        # clamp is active, boundary included (in <= 0 or in >= size-1).
        py = torch.where((py > 0) & (py < h - 1), py, py.detach().clamp(0, h - 1))
        px = torch.where((px > 0) & (px < w - 1), px, px.detach().clamp(0, w - 1))
    y0 = torch.floor(py)
    x0 = torch.floor(px)
    ly = py - y0
    lx = px - x0
    y0 = y0.long()
    x0 = x0.long()
    flat = x.reshape(n, c, h * w)
    out = x.new_zeros(n, c, py.shape[1])
    for dy, wy in ((0, 1 - ly), (1, ly)):
        for dx, wx in ((0, 1 - lx), (1, lx)):
            yy = y0 + dy
            xx = x0 + dx
            ok = (yy >= 0) & (yy <= h - 1) & (xx >= 0) & (xx <= w - 1)
            idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).unsqueeze(1).expand(n, c, -1)
            v = torch.gather(flat, 2, idx)
            wgt = (wy * wx) * ok.to(x.dtype)
            out = out + v * wgt.unsqueeze(1)
    if dcn_reject:
        keep = (py > -1) & (py < h) & (px > -1) & (px < w)
        out = out * keep.to(x.dtype).unsqueeze(1)
    return out


def modulated_deform_conv2d(x, offset, mask, weight, bias=None, stride=1, padding=0,
                            dilation=1, groups=1, deform_groups=1):
    """DCNv2 forward, argument order of mmcv.ops.modulated_deform_conv2d as the
    reference calls it (models/networks.py:627-630).

    out[n,o,y,x] = b[o] + sum_{c,k} W[o,c,k] * mask[n,g(c)*K+k,y,x]
                   * bilin(x[n,c], y*sh-ph+i_k*dh+dy, x*sw-pw+j_k*dw+dx)
    with k=i*kw+j, g(c)=c//(C_in/dg), dy=offset[n,(g*K+k)*2], dx=offset[n,(g*K+k)*2+1].
    """
    sh, sw = _pair(stride)
    ph, pw = _pair(padding)
    dh, dw = _pair(dilation)
    n, cin, h, w = x.shape
    cout, cin_g, kh, kw = weight.shape
    K = kh * kw
    ho = (h + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    wo = (w + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    dg = deform_groups
    assert cin % dg == 0 and cin_g * groups == cin and cout % groups == 0
    assert offset.shape == (n, dg * 2 * K, ho, wo), offset.shape
    assert mask.shape == (n, dg * K, ho, wo), mask.shape
    cpg = cin // dg
    P = ho * wo
    ys = (torch.arange(ho, dtype=x.dtype) * sh - ph).view(ho, 1).expand(ho, wo).reshape(1, P)
    xs = (torch.arange(wo, dtype=x.dtype) * sw - pw).view(1, wo).expand(ho, wo).reshape(1, P)
    off = offset.reshape(n, dg, K, 2, P)
    msk = mask.reshape(n, dg, K, P)
    cols = x.new_zeros(n, cin, K, P)
    for g in range(dg):
        xg = x[:, g * cpg:(g + 1) * cpg]
        for k in range(K):
            i, j = divmod(k, kw)
            py = ys + i * dh + off[:, g, k, 0]
            px = xs + j * dw + off[:, g, k, 1]
            v = bilinear_gather(xg, py, px, dcn_reject=True)
            cols[:, g * cpg:(g + 1) * cpg, k] = v * msk[:, g, k].unsqueeze(1)
    cols = cols.reshape(n, groups, cin_g * K, P)
    wmat = weight.reshape(groups, cout // groups, cin_g * K)
    out = torch.einsum("gok,ngkp->ngop", wmat, cols).reshape(n, cout, ho, wo)
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return out


def flow_warp(x, flow, flow_layout="n2hw", padding_mode="zeros"):
    """flow_warp with align_corners=True, bilinear.

    ``flow_layout='n2hw'``: models/networks.py:699-739 (flow (n,2,h,w)),
    ``flow_layout='nhw2'``: models/eavsrp_model.py:587-626 (flow (n,h,w,2)).
    The reference normalises (x+fx) to [-1,1] with 2/(w-1) and grid_sample
    (align_corners=True) maps it back with (g+1)/2*(w-1): the sample point is
    exactly (y+flow_y, x+flow_x) in pixel units, flow channel 0 = x.
    """
    n, c, h, w = x.shape
    if flow_layout == "n2hw":
        if tuple(flow.shape[-2:]) != (h, w):
            raise ValueError("The spatial sizes of input and flow are not the same.")
        fx, fy = flow[:, 0], flow[:, 1]
    elif flow_layout == "nhw2":
        if tuple(flow.shape[1:3]) != (h, w):
            raise ValueError("The spatial sizes of input and flow are not the same.")
        fx, fy = flow[..., 0], flow[..., 1]
    else:
        raise ValueError(flow_layout)
    # normalisation by 2/max(size-1, 1) (networks.py:730-731) and grid_sample's (g+1)/2*(size-1): along a
    # size-1 dimension the sample coordinate collapses to 0 whatever the flow is
    if w == 1:
        fx = fx * 0
    if h == 1:
        fy = fy * 0
    gy = torch.arange(h, dtype=x.dtype).view(1, h, 1)
    gx = torch.arange(w, dtype=x.dtype).view(1, 1, w)
    py = (gy + fy).reshape(n, h * w)
    px = (gx + fx).reshape(n, h * w)
    out = bilinear_gather(x, py, px, border=(padding_mode == "border"))
    return out.reshape(n, c, h, w)


def backwarp(x, flow):
    """BaseModel.backwarp / get_backwarp (models/base_model.py:321-354) and
    PWCNET.Decoder.backwarp (models/pwc_net.py:184-207): grid_sample with
    align_corners=False on cat[x, ones]; grid = linspace(-1+1/W, 1-1/W) +
    flow/((W-1)/2).  Un-normalising with ((g+1)*W-1)/2 gives the sample point
    x + fx*W/(W-1).  Returns (warped*mask, mask), mask = (warped_ones>0.999).
    """
    n, c, h, w = x.shape
    fx, fy = flow[:, 0], flow[:, 1]
    gy = torch.arange(h, dtype=x.dtype).view(1, h, 1)
    gx = torch.arange(w, dtype=x.dtype).view(1, 1, w)
    py = (gy + fy * (h / (h - 1.0))).reshape(n, h * w)
    px = (gx + fx * (w / (w - 1.0))).reshape(n, h * w)
    xin = torch.cat([x, x.new_ones(n, 1, h, w)], 1)
    out = bilinear_gather(xin, py, px).reshape(n, c + 1, h, w)
    m = (out[:, -1:] > 0.999).to(x.dtype)
    return out[:, :-1] * m, m


def correlation(first, second, max_disp: int = 4):
    """pwc/correlation/correlation.py:35-103: 81-channel cost volume,
    out[n,(dy+4)*9+(dx+4),y,x] = mean_c f1[n,c,y,x]*f2[n,c,y+dy,x+dx], zero outside."""
    n, c, h, w = first.shape
    d = max_disp
    pad = torch.nn.functional.pad(second, (d, d, d, d))
    outs = []
    for iy in range(2 * d + 1):
        for ix in range(2 * d + 1):
            outs.append((first * pad[:, :, iy:iy + h, ix:ix + w]).sum(1) / c)
    return torch.stack(outs, 1)


def correlation_backward(first, second, gout, max_disp: int = 4):
    """Explicit backward per correlation.py:105-233 (not autograd):
    g1[n,c,y,x] = 1/C sum_{dy,dx} gout[n,k,y,x] * f2[n,c,y+dy,x+dx]
    g2[n,c,y,x] = 1/C sum_{dy,dx} gout[n,k,y-dy,x-dx] * f1[n,c,y-dy,x-dx]."""
    n, c, h, w = first.shape
    d = max_disp
    pad2 = torch.nn.functional.pad(second, (d, d, d, d))
    g1 = torch.zeros_like(first)
    g2p = first.new_zeros(n, c, h + 2 * d, w + 2 * d)
    k = 0
    for iy in range(2 * d + 1):
        for ix in range(2 * d + 1):
            go = gout[:, k:k + 1]
            g1 = g1 + go * pad2[:, :, iy:iy + h, ix:ix + w]
            g2p[:, :, iy:iy + h, ix:ix + w] += go * first
            k += 1
    return g1 / c, g2p[:, :, d:d + h, d:d + w] / c


def affine_offsets(transform, translation, deform_groups: int):
    """AdaptBlockOffset's offset expansion (models/networks.py:302-310).

    transform (n,4D,h,w): per group a 2x2 matrix [[a,b],[c,d]] in channel order
    (a,b,c,d); translation (n,2D,h,w).  R = [[-1,-1,-1,0,0,0,1,1,1],
    [-1,0,1,-1,0,1,-1,0,1]] (row 0 = dy of tap k, row 1 = dx).
    offset[n, g*18+2k]   = a*R0k + b*R1k - R0k + t0
    offset[n, g*18+2k+1] = c*R0k + d*R1k - R1k + t1
    """
    n, _, h, w = transform.shape
    D = deform_groups
    R = transform.new_tensor([[-1, -1, -1, 0, 0, 0, 1, 1, 1],
                              [-1, 0, 1, -1, 0, 1, -1, 0, 1]])
    T = transform.reshape(n, D, 2, 2, h, w)
    t = translation.reshape(n, D, 2, h, w)
    off = torch.einsum("ndijhw,jk->ndkihw", T, R) - R.t().reshape(1, 1, 9, 2, 1, 1)
    off = off + t.unsqueeze(2)
    return off.reshape(n, D * 18, h, w)
