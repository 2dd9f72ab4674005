"""Tiny CPU emulator for the reference's CUDA-C kernel *strings* (fixture generation only).

pwc/correlation/correlation.py keeps its four kernels as CUDA-C source strings that cupy
JIT-compiles on a GPU.  Neither cupy nor a GPU exists in the build container, so to obtain
golden vectors from the reference's *own* kernel source we compile the (size-substituted)
string with g++ against a prelude that maps the CUDA execution model onto host threads:
one std::thread per CUDA thread of a block, blocks executed one after another,
run as cooperative fibres -- exactly one runs at a time, in thread-index order, and the baton is
passed at ``__syncthreads()`` / thread exit.  That reproduces the warp-lockstep order the
reference relies on (kernel_Correlation_updateOutput re-zeroes ``sum[]`` right after thread 0 read
it, with no barrier in between: safe on a 32-thread block = one warp, a race for free-running host
threads).  ``__shared__`` = static storage.  Only what those four
kernels use is supported.  Nothing here is copied from the reference: the kernel text is read
from /root/reference at generation time (make_golden.py) and never stored in this repo.
"""
from __future__ import annotations

import ctypes
import hashlib
import re
import subprocess
import tempfile
from pathlib import Path

_PRELUDE = r"""
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
struct uint3_ { unsigned x, y, z; };
static thread_local uint3_ threadIdx;
static uint3_ blockIdx, blockDim, gridDim;
static char* g_dyn_smem = nullptr;
static std::mutex g_m;
static std::condition_variable g_cv;
static unsigned g_turn = 0, g_n = 0;
static std::vector<char> g_done;
static void emu_wait_turn(unsigned t) {
  std::unique_lock<std::mutex> lk(g_m);
  g_cv.wait(lk, [&] { return g_turn == t; });
}
static void emu_pass(unsigned t) {
  {
    std::lock_guard<std::mutex> lk(g_m);
    unsigned nxt = t;
    do { nxt = (nxt + 1) % g_n; } while (g_done[nxt] && nxt != t);
    g_turn = nxt;
  }
  g_cv.notify_all();
}
static void emu_sync() { unsigned t = threadIdx.x; emu_pass(t); emu_wait_turn(t); }
#define __global__
#define __shared__ static
#define __syncthreads() emu_sync()
using std::max;
using std::min;
"""

_RUNNER = r"""
extern "C" void emu_%(name)s(unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned shared_bytes, void** a) {
  std::vector<char> dyn(shared_bytes + 16);
  g_dyn_smem = dyn.data();
  blockDim = {bx, 1, 1};
  gridDim = {gx, gy, gz};
  for (unsigned z = 0; z < gz; ++z) for (unsigned y = 0; y < gy; ++y) for (unsigned x = 0; x < gx; ++x) {
    blockIdx = {x, y, z};
    g_n = bx;
    g_turn = 0;
    g_done.assign(bx, 0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < bx; ++t)
      th.emplace_back([=]() {
        threadIdx = {t, 0, 0};
        emu_wait_turn(t);
        %(name)s(%(call)s);
        { std::lock_guard<std::mutex> lk(g_m); g_done[t] = 1; }
        emu_pass(t);
      });
    for (auto& t : th) t.join();
  }
}
"""


def _parse_params(src: str, name: str):
    m = re.search(name + r"\s*\((.*?)\)\s*\{", src, re.S)
    params = [p.strip() for p in m.group(1).split(",") if p.strip()]
    out = []
    for p in params:
        typ = p.rsplit(None, 1)[0] if "*" not in p else p[:p.rindex("*") + 1]
        out.append(typ.strip())
    return out


def compile_kernel(src: str, name: str):
    src = src.replace("extern __shared__ char patch_data_char[];", "char* patch_data_char = g_dyn_smem;")
    types = _parse_params(src, name)
    call = ", ".join(f"({t})a[{i}]" if "*" in t else f"*({t}*)a[{i}]" for i, t in enumerate(types))
    code = _PRELUDE + src + _RUNNER % {"name": name, "call": call}
    tag = hashlib.sha1(code.encode()).hexdigest()[:16]
    d = Path(tempfile.gettempdir()) / "eavsr_cuda_emu"
    d.mkdir(exist_ok=True)
    so = d / f"{name}_{tag}.so"
    if not so.exists():
        cpp = d / f"{name}_{tag}.cpp"
        cpp.write_text(code)
        subprocess.run(["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread", str(cpp), "-o", str(so)],
                       check=True)
    lib = ctypes.CDLL(str(so))
    fn = getattr(lib, "emu_" + name)
    fn.restype = None
    return fn, types


def launch(src: str, name: str, grid, block, args, shared_mem: int = 0):
    """Run kernel ``name`` of CUDA-C source ``src`` on the CPU.  ``args``: ints, CPU tensors or None."""
    fn, types = compile_kernel(src, name)
    assert len(types) == len(args), (types, args)
    keep, ptrs = [], []
    for t, a in zip(types, args):
        if "*" in t:
            ptrs.append(ctypes.c_void_p(0 if a is None else a.data_ptr()))
        else:
            v = ctypes.c_int(int(a))
            keep.append(v)
            ptrs.append(ctypes.cast(ctypes.pointer(v), ctypes.c_void_p))
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    gx, gy, gz = (list(grid) + [1, 1])[:3]
    assert len(block) == 1 or all(b == 1 for b in block[1:])
    fn(ctypes.c_uint(gx), ctypes.c_uint(gy), ctypes.c_uint(gz), ctypes.c_uint(block[0]),
       ctypes.c_uint(shared_mem), arr)
