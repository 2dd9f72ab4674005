"""Host-side logic of the round-2 operators that needs no GPU: fallbacks on CPU tensors, eligibility predicates and
the stride normalisation at the C boundary."""
import torch

from eavsr_b200 import ops


def test_cat_channels_falls_back_to_torch_cat_on_cpu():
    a, b = torch.randn(2, 8, 5, 7), torch.randn(2, 16, 5, 7)
    assert torch.equal(ops.cat_channels([a, b]), torch.cat([a, b], 1))
    buf = torch.zeros(2, 32, 5, 7)
    out = ops.cat_channels([a, b], out=buf, channel_offset=8)
    assert out is buf and torch.equal(buf[:, 8:], torch.cat([a, b], 1)) and bool((buf[:, :8] == 0).all())


def test_training_helpers_fall_back_on_cpu_and_stay_differentiable():
    conv = torch.nn.Conv2d(64, 64, 3, 1, 1)
    x = torch.randn(1, 64, 6, 6, requires_grad=True)
    y = ops.conv2d_native_bias_grad(conv, x)
    assert torch.allclose(y, conv(x))
    m = ops.channel_mean(x)
    assert torch.allclose(m, x.mean((2, 3), keepdim=True))
    s = torch.rand(1, 64, 1, 1, requires_grad=True)
    z = ops.scale_residual(y, s, x)
    assert torch.allclose(z, y * s + x)
    z.sum().backward()
    assert x.grad is not None and s.grad is not None and conv.bias.grad is not None
    g = torch.nn.Conv2d(128, 64, 3, 1, 1, groups=64)
    assert not ops.grouped_conv3x3_eligible(g, torch.randn(1, 128, 4, 4))      # CPU tensor: nn.Conv2d runs instead


def test_grouped_conv_eligibility_predicate():
    class FakeCuda(torch.Tensor):
        is_cuda = True

    x = torch.randn(1, 128, 4, 4).as_subclass(FakeCuda)
    ok = lambda conv: ops.grouped_conv3x3_eligible(conv, x)      # noqa: E731
    assert ok(torch.nn.Conv2d(128, 128, 3, 1, 1, groups=128)) and ok(torch.nn.Conv2d(128, 64, 3, 1, 1, groups=64))
    assert not ok(torch.nn.Conv2d(128, 32, 3, 1, 1, groups=32))            # 4 inputs per group
    assert not ok(torch.nn.Conv2d(128, 128, 3, 2, 1, groups=128))          # stride 2
    assert not ok(torch.nn.Conv2d(128, 128, 5, 1, 2, groups=128))          # 5x5
    assert not ok(torch.nn.Conv2d(128, 64, 3, 1, 1, groups=32))            # groups != out channels


def test_size_one_batch_stride_is_normalised_at_the_boundary():
    f = torch.randn(5, 64, 17, 24).contiguous(memory_format=torch.channels_last)
    per_frame = f.view(5, 1, 64, 17, 24).unbind(0)[1]            # stride(0) of the size-1 batch is arbitrary (64 here)
    s = ops._strides(per_frame)
    assert list(s)[0] == 64 * 17 * 24 and list(s)[1:] == [1, 24 * 64, 64]
    wide = torch.randn(1, 192, 6, 6).contiguous(memory_format=torch.channels_last)[:, :64]     # channel slice
    assert list(ops._strides(wide))[0] == 6 * 6 * 192
    two = torch.randn(2, 64, 6, 6).contiguous(memory_format=torch.channels_last)
    assert list(ops._strides(two)) == list(two.stride())
