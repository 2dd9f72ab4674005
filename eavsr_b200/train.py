"""Training step of EAVSR+ on the B200 alignment kernels (BASELINE config 5).

Reference: ``EAVSRPModel`` -- optimizer construction models/eavsrp_model.py:45-59 (Adam, two parameter
groups: every ``deform_align.*`` parameter at lr = 1e-5, the rest at ``opt.lr``), ``forward`` :82-98 (incl.
the ``epoch >= npost`` branch that masks the SR with the PWC-Net validity mask of ``get_backwarp``),
``backward`` :109-113 (L1 of the whole sequence, mean), ``optimize_parameters`` :115-119; the loop that calls
it, train_basic.py:58-59.  The reference's multi-GPU mechanism is ``nn.DataParallel`` (models/networks.py:
67-74: one process, batch split, gradients reduced to GPU 0); here it is one process per GPU with
``torch.distributed`` / NCCL ``DistributedDataParallel``: the gradient all-reduce (12.28 M fp32 = 49.1 MB per
step) is the only collective, bucketed and overlapped with the backward pass.  SPyNet is frozen (models/
eavsrp_model.py:132-133), PWC-Net is frozen and evaluated under ``no_grad`` (models/base_model.py:356-360): both
are outside the all-reduce.

Gradients flow through the library's differentiable operators (``flow_warp`` wrt x and flow, DCNv2 wrt x,
offset, mask, weight, bias); the fused inference kernels are disabled whenever autograd is on
(``ops.fused_inference_ok``).  Precision: ``dtype=torch.float32`` is the reference's own (fp32 everywhere);
``torch.bfloat16`` keeps fp32 master weights and runs the network under ``torch.autocast`` so that the
features reach the bf16 tensor-core kernels, flows / offsets / masks stay fp32.
"""
from __future__ import annotations

import contextlib
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["param_groups", "build_optimizer", "Trainer", "trainable_bytes"]

_BRANCHES = ("backward_1", "forward_1", "backward_2", "forward_2")


def param_groups(net: nn.Module, lr: float, align_lr: float = 1e-5) -> List[dict]:
    """The two Adam groups of models/eavsrp_model.py:45-56: ``deform_align[*]`` parameters at ``align_lr``,
    everything else (frozen parameters included, as in the reference -- Adam skips those without a gradient)
    at ``lr``."""
    align_ids = set()
    for b in _BRANCHES:
        align_ids.update(id(p) for p in net.deform_align[b].parameters())
    basic = [p for p in net.parameters() if id(p) not in align_ids]
    align = [p for p in net.parameters() if id(p) in align_ids]
    return [{"params": basic}, {"params": align, "lr": align_lr}]


def build_optimizer(net: nn.Module, lr: float = 1e-4, betas: Tuple[float, float] = (0.9, 0.999),
                    weight_decay: float = 0.0, align_lr: float = 1e-5, capturable: bool = False) -> torch.optim.Adam:
    return torch.optim.Adam(param_groups(net, lr, align_lr), lr=lr, betas=betas, weight_decay=weight_decay,
                            capturable=capturable)


def trainable_bytes(net: nn.Module) -> int:
    """Bytes of gradient the data-parallel all-reduce moves per step (fp32)."""
    return sum(p.numel() for p in net.parameters() if p.requires_grad) * 4


class Trainer:
    """``optimize_parameters`` of the reference (models/eavsrp_model.py:82-119) for one process / one GPU,
    optionally data-parallel over ``torch.distributed``.

    net        : eavsr_b200.model.EAVSRP (or any module with a ``deform_align`` ModuleDict of the 4 branches)
    dtype      : torch.float32 (reference) or torch.bfloat16 (autocast, fp32 master weights)
    ddp        : wrap in DistributedDataParallel (needs an initialised process group)
    pwcnet     : optional eavsr_b200.pwc.PWCNET for the ``epoch >= npost`` branch
    """

    def __init__(self, net: nn.Module, lr: float = 1e-4, betas=(0.9, 0.999), weight_decay: float = 0.0,
                 align_lr: float = 1e-5, dtype: torch.dtype = torch.float32, ddp: bool = False, scale: int = 4,
                 pwcnet: Optional[nn.Module] = None, npost: int = 350, bucket_cap_mb: int = 25,
                 device_ids: Optional[list] = None, capturable: bool = False):
        self.net = net
        self.dtype, self.scale, self.npost = dtype, scale, npost
        self.optimizer = build_optimizer(net, lr, betas, weight_decay, align_lr, capturable)
        self._graph = None
        self.pwcnet = pwcnet
        if pwcnet is not None:
            for p in pwcnet.parameters():
                p.requires_grad = False
        self.model: nn.Module = net
        if ddp:
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("Trainer(ddp=True) needs torch.distributed.init_process_group first")
            from torch.nn.parallel import DistributedDataParallel
            # constructed on a side stream so that the autograd hooks DDP stashes live on the stream the (optionally
            # graph-captured) steps run on -- PyTorch's recipe for DDP under CUDA graphs
            side = torch.cuda.Stream() if torch.cuda.is_available() and capturable else None
            ctx = torch.cuda.stream(side) if side is not None else contextlib.nullcontext()
            if side is not None:
                side.wait_stream(torch.cuda.current_stream())
            with ctx:
                self.model = DistributedDataParallel(net, device_ids=device_ids, bucket_cap_mb=bucket_cap_mb,
                                                     gradient_as_bucket_view=True, broadcast_buffers=False)
            if side is not None:
                torch.cuda.current_stream().wait_stream(side)
            self._ddp_stream = side
        self.loss = None

    # -- pieces of optimize_parameters, exposed for tests and for the bench's overlap measurement --------------
    def _autocast(self):
        if self.dtype == torch.float32:
            return contextlib.nullcontext()
        dev = next(self.net.parameters()).device.type
        return torch.autocast(dev, dtype=self.dtype)

    def forward(self, lr_seq: torch.Tensor, hr_seq: Optional[torch.Tensor] = None, epoch: int = 0) -> torch.Tensor:
        """models/eavsrp_model.py:82-98."""
        with self._autocast():
            sr = self.model(lr_seq)
        sr = sr.float()
        if epoch >= self.npost and self.pwcnet is not None and hr_seq is not None:
            from . import pwc
            masks = [pwc.get_backwarp(lr_seq[:, i].float(), hr_seq[:, i].float(), self.pwcnet, scale=self.scale)[1]
                     for i in range(hr_seq.shape[1])]
            sr = sr * torch.stack(masks, 1)
        return sr

    def compute_loss(self, sr: torch.Tensor, hr_seq: torch.Tensor) -> torch.Tensor:
        """models/eavsrp_model.py:109-111: L1Loss (models/losses.py:9) of the sequences, mean."""
        return F.l1_loss(sr, hr_seq.float())

    def step(self, lr_seq: torch.Tensor, hr_seq: torch.Tensor, epoch: int = 0, sync: bool = True) -> torch.Tensor:
        """forward, zero_grad, backward, optimizer step (models/eavsrp_model.py:115-119).  ``sync=False`` skips the
        gradient all-reduce of this step (DDP ``no_sync``; used to measure what the collective costs)."""
        ctx = contextlib.nullcontext()
        if not sync and hasattr(self.model, "no_sync"):
            ctx = self.model.no_sync()
        with ctx:
            sr = self.forward(lr_seq, hr_seq, epoch)
            self.optimizer.zero_grad(set_to_none=True)
            self.loss = self.compute_loss(sr, hr_seq)
            self.loss.backward()
        self.optimizer.step()
        return self.loss.detach()

    # -- the whole step as one CUDA graph ------------------------------------------------------------------------
    def capture(self, lr_seq: torch.Tensor, hr_seq: torch.Tensor, epoch: int = 0, warmup: int = 3) -> "Trainer":
        """Capture forward + loss + backward + Adam of `step` into ONE CUDA graph (the eager step issues ~45 k launches
        from one Python thread and is bound by the host, not the GPU).  Needs ``Trainer(capturable=True)``; runs
        `warmup` REAL steps on a side stream first (lazy initialisation, cuDNN algorithm selection, DDP bucket
        set-up), then records one step on static copies of the inputs.  `step_graphed` replays it."""
        if not any(g.get("capturable") for g in self.optimizer.param_groups):
            raise RuntimeError("Trainer.capture needs Trainer(..., capturable=True) (Adam state on the device)")
        self._static = (lr_seq.clone(), hr_seq.clone())
        cur = torch.cuda.current_stream()
        side = getattr(self, "_ddp_stream", None) or torch.cuda.Stream()
        side.wait_stream(cur)
        if hasattr(self.model, "no_sync"):
            warmup = max(warmup, 11)            # DDP rebuilds its buckets during the first iterations
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self.step(self._static[0], self._static[1], epoch)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.optimizer.zero_grad(set_to_none=True)
        from . import _lib
        self._graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(self._graph):
            sr = self.forward(self._static[0], self._static[1], epoch)
            self.loss = self.compute_loss(sr, self._static[1])
            self.loss.backward()
            self.optimizer.step()
        self.captured_launches = _lib.launch_count() - l0      # library kernels replayed by every step_graphed
        return self

    def step_graphed(self, lr_seq: torch.Tensor, hr_seq: torch.Tensor) -> torch.Tensor:
        """One training step by graph replay (same shapes as at `capture`)."""
        if self._graph is None:
            raise RuntimeError("call Trainer.capture first")
        self._static[0].copy_(lr_seq, non_blocking=True)
        self._static[1].copy_(hr_seq, non_blocking=True)
        self._graph.replay()
        return self.loss.detach()

    def gradients(self) -> Dict[str, torch.Tensor]:
        return {n: p.grad for n, p in self.net.named_parameters() if p.grad is not None}
