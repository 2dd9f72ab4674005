"""Make the *unmodified* reference pick up the B200 kernels (the drop-in).

    import eavsr_b200.install as I
    I.install()                      # before `import models...` of HITRainer/EAVSR
    ...                              # build EAVSRPModel / EAVSRPx2Model as usual
    I.rebind_flow_warp()             # after the reference modules are imported

The reference resolves its hot path through three import-time symbols:
  * ``from mmcv.ops import ModulatedDeformConv2d, modulated_deform_conv2d`` (models/networks.py:573)
  * module globals ``flow_warp`` of models.networks / models.eavsrp_model / models.eavsrpx2_model
    (looked up at call time, so rebinding the global is enough)
  * ``from pwc.correlation import correlation`` (models/pwc_net.py:19)
``install()`` registers replacement modules in ``sys.modules`` for ``mmcv.ops`` and
``pwc.correlation.correlation``.  When mmcv / cupy themselves are not installed it also registers
the three tiny non-hot-path helpers of mmcv that the reference imports (``mmcv.utils.get_logger``,
``mmcv.runner.load_checkpoint``, ``mmcv.cnn.ConvModule``) and an import-only ``cupy``.
"""
from __future__ import annotations

import importlib
import importlib.util
import logging
import os
import sys
import types

import torch
import torch.nn as nn

from . import ops

__all__ = ["install", "rebind_flow_warp", "rebind_backwarp", "uninstall"]

_installed: dict = {}


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__eavsr_b200__ = True
    return m


class _ConvModule(nn.Module):
    """mmcv.cnn.ConvModule as SPyNet uses it (models/eavsrp_model.py:534-574): conv + optional ReLU,
    children named ``conv`` / ``activate``."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, norm_cfg=None, act_cfg=None,
                 **kw):
        super().__init__()
        if norm_cfg is not None:
            raise NotImplementedError("eavsr_b200 ConvModule shim: norm layers are not used by EAVSR")
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding)
        self.activate = nn.ReLU(inplace=True) if act_cfg is not None else None

    def forward(self, x):
        x = self.conv(x)
        return self.activate(x) if self.activate is not None else x


def _load_checkpoint(model, filename, map_location=None, strict=False, logger=None):
    if not os.path.exists(filename):
        raise FileNotFoundError(filename)
    sd = torch.load(filename, map_location=map_location or "cpu")
    sd = sd.get("state_dict", sd)
    model.load_state_dict(sd, strict=strict)
    return sd


def _get_logger(name, log_file=None, log_level=logging.INFO):
    logger = logging.getLogger(name)
    logger.setLevel(log_level)
    return logger


def install(force_shims: bool = False) -> None:
    """Register the replacement modules.  Idempotent."""
    def put(name, mod):
        _installed.setdefault(name, sys.modules.get(name))
        sys.modules[name] = mod

    have_mmcv = importlib.util.find_spec("mmcv") is not None and not force_shims
    mmcv_ops = _module("mmcv.ops", ModulatedDeformConv2d=ops.ModulatedDeformConv2d,
                       modulated_deform_conv2d=ops.modulated_deform_conv2d)
    if have_mmcv:
        import mmcv  # noqa: F401
        put("mmcv.ops", mmcv_ops)
        sys.modules["mmcv"].ops = mmcv_ops
    else:
        root = _module("mmcv", __path__=[])
        root.ops = mmcv_ops
        root.utils = _module("mmcv.utils", get_logger=_get_logger)
        root.runner = _module("mmcv.runner", load_checkpoint=_load_checkpoint)
        root.cnn = _module("mmcv.cnn", ConvModule=_ConvModule)
        for name, mod in (("mmcv", root), ("mmcv.ops", mmcv_ops), ("mmcv.utils", root.utils),
                          ("mmcv.runner", root.runner), ("mmcv.cnn", root.cnn)):
            put(name, mod)
    corr = _module("pwc.correlation.correlation", FunctionCorrelation=ops.FunctionCorrelation,
                   ModuleCorrelation=ops.ModuleCorrelation, _FunctionCorrelation=ops._FunctionCorrelation)
    pkg = sys.modules.get("pwc.correlation")
    if pkg is None:
        try:
            pkg = importlib.import_module("pwc.correlation")
        except Exception:                    # reference not on sys.path yet: provide the package too
            put("pwc", _module("pwc", __path__=[]))
            pkg = _module("pwc.correlation", __path__=[])
            put("pwc.correlation", pkg)
    pkg.correlation = corr
    put("pwc.correlation.correlation", corr)


def rebind_flow_warp() -> list:
    """Point the reference's module-level ``flow_warp`` globals at the CUDA op.  Returns the list of
    module names that were patched (those already imported)."""
    done = []
    for name, fn in (("models.networks", ops.flow_warp), ("models.eavsrp_model", ops.flow_warp_nhw2),
                     ("models.eavsrpx2_model", ops.flow_warp_nhw2)):
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "flow_warp"):
            mod.flow_warp = fn
            done.append(name)
    return done


def rebind_backwarp(model=None) -> list:
    """Point the reference's train-time backwarp helpers at the CUDA op:
    ``BaseModel.get_backwarp`` (models/base_model.py:344-354; the flow-estimation branch is kept) on the
    class, and -- because ``Decoder`` is a class local to ``PWCNET.__init__`` (models/pwc_net.py:97-207) --
    the bound ``backwarp`` of every Decoder instance below ``model``.  Returns what was patched."""
    done = []
    bm = sys.modules.get("models.base_model")
    if bm is not None and hasattr(bm, "BaseModel"):
        import torch.nn.functional as F

        def get_backwarp(self, tenFirst, tenSecond, net, flow=None, scale=1):
            if flow is None:
                second = F.interpolate(tenSecond, scale_factor=1 / scale, mode='bilinear', align_corners=True)
                flow = self.get_flow(tenFirst, second, net)
                flow = F.interpolate(flow, scale_factor=scale, mode='nearest') * scale
            return ops.get_backwarp(tenSecond, flow)

        bm.BaseModel.get_backwarp = get_backwarp
        done.append("models.base_model.BaseModel.get_backwarp")
    if model is not None:
        for name, m in model.named_modules():
            if type(m).__name__ == "Decoder" and hasattr(m, "backwarp"):
                m.backwarp = ops.backwarp
                done.append(f"{name}.backwarp")
    return done


def uninstall() -> None:
    for name, old in _installed.items():
        if old is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = old
    _installed.clear()
