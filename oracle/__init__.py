"""CPU oracle for the EAVSR alignment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``eavsr_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the reported CPU baseline -- never as the thing shipped.

Pinning status (see DESIGN.md "Oracle"):
  * flow_warp (both flow layouts, zeros/border) and BaseModel.backwarp:
    pinned against the reference's own functions imported from /root/reference
    (fixtures in tests/golden/, generator tests/golden/make_golden.py).
  * DCNv2: the arithmetic lives in mmcv (mmcv-full 1.x, un-pinned, NOT under
    /root/reference).  Pinned against torchvision.ops.deform_conv2d (same MSRA
    DCNv2 lineage) through the reference's own call site
    (models/networks.py:627-630) with an mmcv.ops shim -> "parity pinned to the
    stand-in, un-pinned to mmcv itself".
  * correlation: the reference's CUDA-C kernel strings
    (pwc/correlation/correlation.py:8-233) cannot run without cupy + a GPU;
    fixtures are produced by executing those very kernel strings under a CPU
    CUDA-emulation harness (tests/golden/cuda_emu.py) at generation time.
"""
