"""CPU tests of the training-step harness (eavsr_b200.train): the two Adam groups of
models/eavsrp_model.py:45-59, the optimize_parameters order of :115-119, and -- world_size 2 over gloo -- that
the data-parallel gradients after the all-reduce equal the single-process gradients on the concatenated
batch.  The real network cannot run on the CPU (the operators have no CPU path), so a stand-in with the same
`deform_align` structure is used for the host logic; the GPU tests run the real one."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from eavsr_b200 import train as T
from eavsr_b200.model import EAVSRP

BR = ("backward_1", "forward_1", "backward_2", "forward_2")


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.deform_align = nn.ModuleDict({b: nn.Conv2d(3, 3, 3, 1, 1) for b in BR})
        self.body = nn.Conv2d(3, 3, 3, 1, 1)
        self.frozen = nn.Conv2d(3, 3, 1)
        for p in self.frozen.parameters():
            p.requires_grad = False

    def forward(self, lrs):
        n, t, c, h, w = lrs.shape
        x = lrs.reshape(n * t, c, h, w)
        x = x + self.frozen(x)
        for b in BR:
            x = x + torch.tanh(self.deform_align[b](x))
        x = nn.functional.interpolate(self.body(x), scale_factor=4, mode="bilinear", align_corners=False)
        return x.view(n, t, c, 4 * h, 4 * w)


def _data(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, 3, 8, 8, generator=g), torch.rand(n, 3, 3, 32, 32, generator=g)


def test_param_groups_follow_the_reference():
    net = EAVSRP(4, n_resblock=1)
    groups = T.param_groups(net, 1e-4)
    align = {id(p) for b in BR for p in net.deform_align[b].parameters()}
    assert {id(p) for p in groups[1]["params"]} == align and groups[1]["lr"] == 1e-5
    assert {id(p) for p in groups[0]["params"]} == {id(p) for p in net.parameters()} - align
    opt = T.build_optimizer(net, lr=2e-4, betas=(0.9, 0.99))
    assert [g["lr"] for g in opt.param_groups] == [2e-4, 1e-5] and opt.param_groups[0]["betas"] == (0.9, 0.99)
    # 12 277 799 trainable parameters -> 49.1 MB of fp32 gradients per step at 30 blocks
    full = EAVSRP(4)
    assert T.trainable_bytes(full) == 12_277_799 * 4


def test_step_is_forward_zero_grad_backward_adam():
    torch.manual_seed(0)
    net = Tiny()
    ref = Tiny()
    ref.load_state_dict(net.state_dict())
    lr, hr = _data(2)
    tr = T.Trainer(net, lr=1e-3)
    loss = tr.step(lr, hr)
    opt = torch.optim.Adam([{"params": [p for n, p in ref.named_parameters() if not n.startswith("deform_align")]},
                            {"params": [p for n, p in ref.named_parameters() if n.startswith("deform_align")], "lr": 1e-5}],
                           lr=1e-3)
    out = ref(lr)
    opt.zero_grad()
    l2 = (out - hr).abs().mean()
    l2.backward()
    opt.step()
    assert torch.allclose(loss, l2.detach())
    for (n, a), (_, b) in zip(net.named_parameters(), ref.named_parameters()):
        assert torch.equal(a, b), n
    assert net.frozen.weight.grad is None


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = Tiny()
        tr = T.Trainer(net, lr=1e-3, ddp=True)
        lr, hr = _data(4)
        half = slice(rank * 2, rank * 2 + 2)
        loss = tr.step(lr[half], hr[half])
        ret[rank] = ({n: g.clone() for n, g in tr.gradients().items()}, {n: p.detach().clone() for n, p in net.named_parameters()},
                     loss.item())
    finally:
        dist.destroy_process_group()


def test_two_rank_ddp_gradients_equal_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    torch.manual_seed(0)
    net = Tiny()
    tr = T.Trainer(net, lr=1e-3)
    lr, hr = _data(4)
    tr.step(lr, hr)
    single = tr.gradients()
    for r in (0, 1):
        grads, params, _ = ret[r]
        assert sorted(grads) == sorted(single)
        for n in single:
            assert (grads[n] - single[n]).abs().max().item() <= 1e-5 * max(1.0, single[n].abs().max().item()), n
        for n, p in net.named_parameters():
            assert torch.allclose(params[n], p.detach(), atol=1e-6), n
    assert abs(0.5 * (ret[0][2] + ret[1][2]) - tr.loss.item()) < 1e-6


def test_ddp_without_process_group_raises():
    with pytest.raises(RuntimeError, match="init_process_group"):
        T.Trainer(Tiny(), ddp=True)
