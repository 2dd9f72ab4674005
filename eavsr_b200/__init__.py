"""eavsr_b200 -- B200-native (sm_100a) inter-frame alignment kernels for EAVSR.

DCNv2 (tcgen05 implicit GEMM), flow_warp and the PWC-Net cost volume behind the reference's own
Python call signatures.  See DESIGN.md / INTEGRATION.md.
"""
from . import _lib  # noqa: F401
from .ops import (  # noqa: F401
    FunctionCorrelation,
    ModuleCorrelation,
    ModulatedDeformConv2d,
    backwarp,
    dcn_affine,
    dcn_affine_eligible,
    dcn_uses_tensor_cores,
    flow_warp,
    flow_warp_nhw2,
    get_backwarp,
    invalidate_caches,
    modulated_deform_conv2d,
)

__version__ = "0.2.0"
