"""Seeded synthetic clips and reference-independent seeded weights (no datasets or checkpoints are
available offline: ckpt/realvsr/*.pth, SPyNet and VGG16 weights are all stripped)."""
import torch


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def seeded_parameters(module):
    """Deterministic, reference-independent weights: every floating-point *parameter* is refilled
    from a generator seeded by crc32(parameter name) -- U(+-1/sqrt(fan_in)) for weights (PyTorch's
    default conv bound, keeps the 30-block residual stacks O(1)), U(+-0.05) for 1-D tensors.
    Buffers (mean/std/regular_matrix) keep their constructor values.  Returns {name: shape}."""
    import math
    import zlib
    shapes = {}
    with torch.no_grad():
        for name, p in sorted(module.named_parameters()):
            g = gen(zlib.crc32(name.encode()) & 0x7FFFFFFF)
            if p.dim() >= 2:
                bound = 1.0 / math.sqrt(max(1, p[0].numel()))
            else:
                bound = 0.05
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
            shapes[name] = tuple(p.shape)
    return shapes


def clip_inputs(n, t, h, w, seed=0):
    """Synthetic LR clip in [0,1]: low-frequency texture translated by a smooth sub-pixel
    trajectory (<= 3 px / frame) plus N(0, 0.01) noise."""
    import math
    g = gen(seed)
    yy = torch.arange(h, dtype=torch.float32).view(1, h, 1)
    xx = torch.arange(w, dtype=torch.float32).view(1, 1, w)
    fr = torch.rand(3, 8, 2, generator=g) * 0.25 + 0.02      # radians / pixel
    ph = torch.rand(3, 8, generator=g) * 2 * math.pi
    am = torch.rand(3, 8, generator=g) / 8
    traj = torch.cumsum((torch.rand(t, 2, generator=g) * 2 - 1) * 3.0, 0)
    frames = []
    for i in range(t):
        dy, dx = traj[i, 0].item(), traj[i, 1].item()
        img = torch.full((3, h, w), 0.5)
        for c in range(3):
            for k in range(8):
                img[c] += am[c, k] * torch.sin(fr[c, k, 0] * (yy[0] + dy) + fr[c, k, 1] * (xx[0] + dx) + ph[c, k])
        frames.append(img)
    clip = torch.stack(frames, 0).unsqueeze(0).repeat(n, 1, 1, 1, 1)
    clip = clip + torch.randn(clip.shape, generator=g) * 0.01
    return clip.clamp(0, 1)
