// PWC-Net cost volume (81 displacements, max displacement 4) for sm_100a.
//
// Replaces the reference's cupy/NVRTC kernels (pwc/correlation/correlation.py):
//   kernel_Correlation_rearrange (:8-33)  -- NCHW -> zero-padded NHWC copies of both inputs,
//   kernel_Correlation_updateOutput (:35-103) -- one 32-thread block per output pixel,
//   kernel_Correlation_updateGradFirst/Second (:105-233) -- launched per sample from Python.
// Here there are no padded temporaries: a CTA stages an 8x32 tile of `first` and the matching
// 16x40 halo tile of `second` (zero-filled outside the image == the reference's zero padding)
// in shared memory, 8 channels at a time, and every thread keeps 4 pixels x 27 displacements
// (3 displacement rows) in registers, so one 16-byte shared load feeds 9-12 FMAs.
// out[n,(dy+4)*9+(dx+4),y,x] = 1/C * sum_c first[n,c,y,x] * second[n,c,y+dy,x+dx].
#include <cuda.h>

#include "common.cuh"

namespace eavsr {
namespace {

constexpr int D = 4;            // max displacement
constexpr int ND = 2 * D + 1;   // 9
constexpr int TH = 8, TW = 32;  // output tile
constexpr int CKC = 8;          // channels per staging pass
constexpr int HH = TH + 2 * D, HW_ = TW + 2 * D;  // 16 x 40 halo tile
constexpr int CORR_THREADS = 8 * TH * 3;          // (x-quad, row, displacement-row group)

template <typename T>
__global__ void __launch_bounds__(CORR_THREADS)
corr_fwd(const T* __restrict__ f1, const T* __restrict__ f2, T* __restrict__ out, int C, int H, int W) {
  __shared__ __align__(16) float s1[CKC][TH][TW];
  __shared__ __align__(16) float s2[CKC][HH][HW_];
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
  const int tid = threadIdx.x;
  const int tq = tid & 7, ty = (tid >> 3) & 7, dgrp = tid >> 6;
  const size_t plane = (size_t)H * W;
  const T* f1n = f1 + (size_t)n * C * plane;
  const T* f2n = f2 + (size_t)n * C * plane;

  float acc[3][ND][4];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < ND; ++b)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[a][b][p] = 0.f;

  for (int c0 = 0; c0 < C; c0 += CKC) {
    __syncthreads();
    for (int i = tid; i < CKC * TH * TW; i += CORR_THREADS) {
      const int xx = i % TW, yy = (i / TW) % TH, cc = i / (TW * TH);
      const int gy = y0 + yy, gx = x0 + xx, gc = c0 + cc;
      float v = 0.f;
      if (gc < C && gy < H && gx < W) v = to_f32<T>(f1n[(size_t)gc * plane + (size_t)gy * W + gx]);
      s1[cc][yy][xx] = v;
    }
    for (int i = tid; i < CKC * HH * HW_; i += CORR_THREADS) {
      const int xx = i % HW_, yy = (i / HW_) % HH, cc = i / (HW_ * HH);
      const int gy = y0 + yy - D, gx = x0 + xx - D, gc = c0 + cc;
      float v = 0.f;
      if (gc < C && gy >= 0 && gy < H && gx >= 0 && gx < W) v = to_f32<T>(f2n[(size_t)gc * plane + (size_t)gy * W + gx]);
      s2[cc][yy][xx] = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int cc = 0; cc < CKC; ++cc) {
      const float4 a4 = *reinterpret_cast<const float4*>(&s1[cc][ty][4 * tq]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int dyi = 0; dyi < 3; ++dyi) {
        const float* row = &s2[cc][ty + dgrp * 3 + dyi][4 * tq];
        const float4 b0 = *reinterpret_cast<const float4*>(row);
        const float4 b1 = *reinterpret_cast<const float4*>(row + 4);
        const float4 b2 = *reinterpret_cast<const float4*>(row + 8);
        const float b[12] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
        for (int dxi = 0; dxi < ND; ++dxi)
#pragma unroll
          for (int p = 0; p < 4; ++p) acc[dyi][dxi][p] += a[p] * b[p + dxi];
      }
    }
  }

  const int gy = y0 + ty, gx = x0 + 4 * tq;
  if (gy >= H || gx >= W) return;
  const float inv = 1.f / (float)C;
  T* on = out + (size_t)n * (ND * ND) * plane + (size_t)gy * W + gx;
  const bool vec = (sizeof(T) == 4) && (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
#pragma unroll
  for (int dyi = 0; dyi < 3; ++dyi)
#pragma unroll
    for (int dxi = 0; dxi < ND; ++dxi) {
      const int k = (dgrp * 3 + dyi) * ND + dxi;
      T* op = on + (size_t)k * plane;
      if (vec) {  // W % 4 == 0 and gx % 4 == 0 -> whole quad in range and 16-byte aligned
        *reinterpret_cast<float4*>(op) = make_float4(acc[dyi][dxi][0] * inv, acc[dyi][dxi][1] * inv,
                                                     acc[dyi][dxi][2] * inv, acc[dyi][dxi][3] * inv);
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if (gx + p < W) op[p] = from_f32<T>(acc[dyi][dxi][p] * inv);
      }
    }
}

// ---- fp32 fast path: cp.async double-buffered staging (zero-fill outside the image) --------------
// The first version staged with scalar loads + stores between two __syncthreads() and reached 7 % of
// HBM on the 30x32x80x128 level (382 us): the FMA work (81*C per pixel, ~25 us at the FP32 pipe's
// peak) was serialised behind the staging.  Here channel chunk k+1 streams into the second buffer
// with 16-byte cp.async (src-size 0 => zero fill for the padding) while chunk k is consumed.
struct CorrStage {
  float s1[CKC][TH][TW];
  float s2[CKC][HH][HW_];
};

__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_4_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}

template <bool VEC4>
__global__ void __launch_bounds__(CORR_THREADS, 2)
corr_fwd_f32(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ out, int C, int H,
             int W) {
  extern __shared__ __align__(16) uint8_t corr_smem[];
  CorrStage* st = reinterpret_cast<CorrStage*>(corr_smem);
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
  const int tid = threadIdx.x;
  const int tq = tid & 7, ty = (tid >> 3) & 7, dgrp = tid >> 6;
  const size_t plane = (size_t)H * W;
  const float* f1n = f1 + (size_t)n * C * plane;
  const float* f2n = f2 + (size_t)n * C * plane;

  auto stage_in = [&](int c0, int buf) {
    CorrStage& S = st[buf];
    if (VEC4) {
      for (int i = tid; i < CKC * TH * (TW / 4); i += CORR_THREADS) {
        const int xx = (i % (TW / 4)) * 4, yy = (i / (TW / 4)) % TH, cc = i / ((TW / 4) * TH);
        const int gy = y0 + yy, gx = x0 + xx, gc = c0 + cc;
        const bool ok = gc < C && gy < H && gx < W;
        cp_async_16_zfill(smem_u32(&S.s1[cc][yy][xx]), ok ? f1n + (size_t)gc * plane + (size_t)gy * W + gx : f1n, ok);
      }
      for (int i = tid; i < CKC * HH * (HW_ / 4); i += CORR_THREADS) {
        const int xx = (i % (HW_ / 4)) * 4, yy = (i / (HW_ / 4)) % HH, cc = i / ((HW_ / 4) * HH);
        const int gy = y0 + yy - D, gx = x0 + xx - D, gc = c0 + cc;
        const bool ok = gc < C && gy >= 0 && gy < H && gx >= 0 && gx < W;
        cp_async_16_zfill(smem_u32(&S.s2[cc][yy][xx]), ok ? f2n + (size_t)gc * plane + (size_t)gy * W + gx : f2n, ok);
      }
    } else {
      for (int i = tid; i < CKC * TH * TW; i += CORR_THREADS) {
        const int xx = i % TW, yy = (i / TW) % TH, cc = i / (TW * TH);
        const int gy = y0 + yy, gx = x0 + xx, gc = c0 + cc;
        const bool ok = gc < C && gy < H && gx < W;
        cp_async_4_zfill(smem_u32(&S.s1[cc][yy][xx]), ok ? f1n + (size_t)gc * plane + (size_t)gy * W + gx : f1n, ok);
      }
      for (int i = tid; i < CKC * HH * HW_; i += CORR_THREADS) {
        const int xx = i % HW_, yy = (i / HW_) % HH, cc = i / (HW_ * HH);
        const int gy = y0 + yy - D, gx = x0 + xx - D, gc = c0 + cc;
        const bool ok = gc < C && gy >= 0 && gy < H && gx >= 0 && gx < W;
        cp_async_4_zfill(smem_u32(&S.s2[cc][yy][xx]), ok ? f2n + (size_t)gc * plane + (size_t)gy * W + gx : f2n, ok);
      }
    }
    cp_async_commit();
  };

  float acc[3][ND][4];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < ND; ++b)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[a][b][p] = 0.f;

  const int nchunks = (C + CKC - 1) / CKC;
  stage_in(0, 0);
  for (int k = 0; k < nchunks; ++k) {
    if (k + 1 < nchunks) {
      stage_in((k + 1) * CKC, (k + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const CorrStage& S = st[k & 1];
#pragma unroll 2
    for (int cc = 0; cc < CKC; ++cc) {
      const float4 a4 = *reinterpret_cast<const float4*>(&S.s1[cc][ty][4 * tq]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int dyi = 0; dyi < 3; ++dyi) {
        const float* row = &S.s2[cc][ty + dgrp * 3 + dyi][4 * tq];
        const float4 b0 = *reinterpret_cast<const float4*>(row);
        const float4 b1 = *reinterpret_cast<const float4*>(row + 4);
        const float4 b2 = *reinterpret_cast<const float4*>(row + 8);
        const float b[12] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
        for (int dxi = 0; dxi < ND; ++dxi)
#pragma unroll
          for (int p = 0; p < 4; ++p) acc[dyi][dxi][p] += a[p] * b[p + dxi];
      }
    }
    __syncthreads();   // everyone is done with buffer k&1 before chunk k+2 streams into it
  }

  const int gy = y0 + ty, gx = x0 + 4 * tq;
  if (gy >= H || gx >= W) return;
  const float inv = 1.f / (float)C;
  float* on = out + (size_t)n * (ND * ND) * plane + (size_t)gy * W + gx;
#pragma unroll
  for (int dyi = 0; dyi < 3; ++dyi)
#pragma unroll
    for (int dxi = 0; dxi < ND; ++dxi) {
      const int kk = (dgrp * 3 + dyi) * ND + dxi;
      float* op = on + (size_t)kk * plane;
      if (VEC4) {
        *reinterpret_cast<float4*>(op) = make_float4(acc[dyi][dxi][0] * inv, acc[dyi][dxi][1] * inv,
                                                     acc[dyi][dxi][2] * inv, acc[dyi][dxi][3] * inv);
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if (gx + p < W) op[p] = acc[dyi][dxi][p] * inv;
      }
    }
}

// ---- fp32 main path: TMA-staged halos, persistent CTAs -------------------------------------------
// ncu of corr_fwd_f32 (profiles/r1_corr_fwd_f32_ncu.txt): 47 M warp instructions of which only 25 M are FMAs --
// the cp.async staging loops (index arithmetic per 16 bytes) cost as much issue bandwidth as a third of
// the math, every CTA exposes the latency of its first chunk, and 1200 CTAs on 296 slots run 5 waves for
// 4.05 waves of work.  Here one elected thread issues two cp.async.bulk.tensor (4-D tensor maps over
// (x, y, c, n); out-of-range coordinates -- the zero padding of the cost volume, the channel tail --
// are filled with zeros by the TMA unit) per 8-channel chunk into a 3-stage ring, the six warps
// only wait on mbarriers, and CTAs are persistent so the ring keeps streaming across tile boundaries.
constexpr int TMA_STAGES = 3;
constexpr int TMA_STAGE_BYTES = (int)sizeof(CorrStage);            // 8 KB + 20 KB
constexpr int TMA_THREADS = CORR_THREADS;

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int x, int y, int c, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(c), "r"(n), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(TMA_THREADS, 2)
corr_fwd_tma(const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2,
             float* __restrict__ out, int C, int H, int W, int tiles_x, int tiles_y, int total_tiles) {
  extern __shared__ __align__(128) uint8_t corr_smem[];
  const uint32_t sbase = (smem_u32(corr_smem) + 127u) & ~127u;
  const uint8_t* sgen = corr_smem + (sbase - smem_u32(corr_smem));
  const uint32_t bars = sbase + TMA_STAGES * TMA_STAGE_BYTES;      // full[TMA_STAGES], empty[TMA_STAGES]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int CWARPS = CORR_THREADS / 32;
  if (tid == 0) {
    for (int s = 0; s < TMA_STAGES; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (TMA_STAGES + s), CWARPS); }
    fence_mbar_init();
  }
  __syncthreads();
  const int nchunks = (C + CKC - 1) / CKC;
  const size_t plane = (size_t)H * W;

  // producer state (thread 0): stream position `pcnt` runs TMA_STAGES-1 chunks ahead of the consumers
  int ptile = blockIdx.x, pk = 0, pcnt = 0;
  auto produce = [&]() {
    if (ptile >= total_tiles) return;
    const int n = ptile / (tiles_x * tiles_y), rem = ptile - n * (tiles_x * tiles_y);
    const int y0 = (rem / tiles_x) * TH, x0 = (rem % tiles_x) * TW;
    const int s = pcnt % TMA_STAGES;
    if (pcnt >= TMA_STAGES) mbar_wait(bars + 8 * (TMA_STAGES + s), ((pcnt / TMA_STAGES) - 1) & 1);
    const uint32_t full = bars + 8 * s, dst = sbase + s * TMA_STAGE_BYTES;
    mbar_arrive_expect_tx(full, TMA_STAGE_BYTES);
    tma_load_4d(dst, &tm1, x0, y0, pk * CKC, n, full);
    tma_load_4d(dst + (uint32_t)sizeof(float) * CKC * TH * TW, &tm2, x0 - D, y0 - D, pk * CKC, n, full);
    ++pcnt;
    if (++pk == nchunks) { pk = 0; ptile += gridDim.x; }
  };
  if (tid == 0)
    for (int i = 0; i < TMA_STAGES - 1; ++i) produce();

  const int tq = tid & 7, ty = (tid >> 3) & 7, dgrp = tid >> 6;
  const float inv = 1.f / (float)C;
  int cnt = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y), rem = tile - n * (tiles_x * tiles_y);
    const int y0 = (rem / tiles_x) * TH, x0 = (rem % tiles_x) * TW;
    float acc[3][ND][4];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < ND; ++b)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[a][b][p] = 0.f;
    for (int k = 0; k < nchunks; ++k, ++cnt) {
      const int s = cnt % TMA_STAGES;
      if (tid == 0) produce();                               // refills the stage everyone left in iteration cnt-1
      mbar_wait(bars + 8 * s, (cnt / TMA_STAGES) & 1);
      const CorrStage& S = *reinterpret_cast<const CorrStage*>(sgen + s * TMA_STAGE_BYTES);
#pragma unroll 2
      for (int cc = 0; cc < CKC; ++cc) {
        const float4 a4 = *reinterpret_cast<const float4*>(&S.s1[cc][ty][4 * tq]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int dyi = 0; dyi < 3; ++dyi) {
          const float* row = &S.s2[cc][ty + dgrp * 3 + dyi][4 * tq];
          const float4 b0 = *reinterpret_cast<const float4*>(row);
          const float4 b1 = *reinterpret_cast<const float4*>(row + 4);
          const float4 b2 = *reinterpret_cast<const float4*>(row + 8);
          const float b[12] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
          for (int dxi = 0; dxi < ND; ++dxi)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[dyi][dxi][p] += a[p] * b[p + dxi];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * (TMA_STAGES + s));
    }
    const int gy = y0 + ty, gx = x0 + 4 * tq;
    if (gy < H && gx < W) {                                  // W % 4 == 0 on this path: whole quad in range
      float* on = out + (size_t)n * (ND * ND) * plane + (size_t)gy * W + gx;
#pragma unroll
      for (int dyi = 0; dyi < 3; ++dyi)
#pragma unroll
        for (int dxi = 0; dxi < ND; ++dxi) {
          const int kk = (dgrp * 3 + dyi) * ND + dxi;
          __stcs(reinterpret_cast<float4*>(on + (size_t)kk * plane),
                 make_float4(acc[dyi][dxi][0] * inv, acc[dyi][dxi][1] * inv, acc[dyi][dxi][2] * inv,
                             acc[dyi][dxi][3] * inv));
        }
    }
  }
}

// ---- small maps (the PWC pyramid as EAVSR training uses it: 1x1 ... 16x16) ------------------------
// One thread per output element, channels in the inner loop; neighbouring threads are neighbouring
// x, so both reads are coalesced and everything lives in L1/L2.  The tiled kernels above would stage
// a whole 8x32 tile (+ halo) per 8 channels for a handful of pixels.
template <typename T>
__global__ void __launch_bounds__(256)
corr_fwd_small(const T* __restrict__ f1, const T* __restrict__ f2, T* __restrict__ out, int N, int C, int H, int W) {
  const int plane = H * W;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * (ND * ND) * plane) return;
  const int x = (int)(idx % W), y = (int)((idx / W) % H);
  const int k = (int)((idx / plane) % (ND * ND)), n = (int)(idx / ((long long)plane * ND * ND));
  const int yy = y + k / ND - D, xx = x + k % ND - D;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
    const T* p1 = f1 + (size_t)n * C * plane + (size_t)y * W + x;
    const T* p2 = f2 + (size_t)n * C * plane + (size_t)yy * W + xx;
    int c = 0;
    for (; c + 4 <= C; c += 4) {
      a0 += to_f32<T>(p1[(size_t)c * plane]) * to_f32<T>(p2[(size_t)c * plane]);
      a1 += to_f32<T>(p1[(size_t)(c + 1) * plane]) * to_f32<T>(p2[(size_t)(c + 1) * plane]);
      a2 += to_f32<T>(p1[(size_t)(c + 2) * plane]) * to_f32<T>(p2[(size_t)(c + 2) * plane]);
      a3 += to_f32<T>(p1[(size_t)(c + 3) * plane]) * to_f32<T>(p2[(size_t)(c + 3) * plane]);
    }
    for (; c < C; ++c) a0 += to_f32<T>(p1[(size_t)c * plane]) * to_f32<T>(p2[(size_t)c * plane]);
  }
  out[idx] = from_f32<T>(((a0 + a1) + (a2 + a3)) / (float)C);
}

// Backward, one thread per input-gradient element; reads are coalesced along x and hit L1/L2.
// gfirst[n,c,y,x]  = 1/C sum_k gout[n,k,y,x]       * second[n,c,y+dy,x+dx]
// gsecond[n,c,y,x] = 1/C sum_k gout[n,k,y-dy,x-dx] * first[n,c,y-dy,x-dx]
template <typename T, bool SECOND>
__global__ void __launch_bounds__(256)
corr_bwd(const T* __restrict__ other, const T* __restrict__ gout, T* __restrict__ gin, int N, int C, int H, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t plane = (size_t)H * W;
  if (idx >= (long long)N * C * plane) return;
  const int x = (int)(idx % W), y = (int)((idx / W) % H);
  const int c = (int)((idx / plane) % C), n = (int)(idx / (plane * C));
  const T* on = other + ((size_t)n * C + c) * plane;
  const T* gn = gout + (size_t)n * (ND * ND) * plane;
  float a = 0.f;
#pragma unroll 1
  for (int dy = -D; dy <= D; ++dy) {
    const int yy = SECOND ? y - dy : y + dy;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int dx = -D; dx <= D; ++dx) {
      const int xx = SECOND ? x - dx : x + dx;
      if (xx < 0 || xx >= W) continue;
      const int k = (dy + D) * ND + (dx + D);
      const float g = SECOND ? to_f32<T>(gn[(size_t)k * plane + (size_t)yy * W + xx])
                             : to_f32<T>(gn[(size_t)k * plane + (size_t)y * W + x]);
      a += g * to_f32<T>(on[(size_t)yy * W + xx]);
    }
  }
  gin[idx] = from_f32<T>(a / (float)C);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// (x, y, c, n) fp32 tensor map with a [box_w x box_h x CKC x 1] box; zero fill outside
bool make_corr_map(CUtensorMap* tm, const void* base, int n, int c, int h, int w, int box_w, int box_h) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)c, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)w * 4, (cuuint64_t)w * h * 4, (cuuint64_t)w * h * c * 4};
  const cuuint32_t box[4] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)CKC, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T>
int corr_forward_t(const void* f1, const void* f2, void* out, int n, int c, int h, int w, cudaStream_t st) {
  if (h * w <= 256) {                       // PWC pyramid of 64x64 training crops: 1x1 ... 16x16
    const long long total = (long long)n * ND * ND * h * w;
    corr_fwd_small<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const T*)f1, (const T*)f2, (T*)out, n, c, h, w);
    return check_launch("correlation_forward(small)");
  }
  dim3 grid(ceil_div(w, TW), ceil_div(h, TH), n);
  if (sizeof(T) == 4) {
    const bool vec = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(f1) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(f2) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    if (vec && (long long)w * h * c * 4 < (1ll << 40)) {
      CUtensorMap tm1, tm2;
      if (make_corr_map(&tm1, f1, n, c, h, w, TW, TH) && make_corr_map(&tm2, f2, n, c, h, w, HW_, HH)) {
        const int smem = TMA_STAGES * TMA_STAGE_BYTES + 2 * TMA_STAGES * 8 + 128;
        cudaError_t e = cudaFuncSetAttribute(corr_fwd_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("correlation_forward: smem attr: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int tiles_x = ceil_div(w, TW), tiles_y = ceil_div(h, TH);
        const long long total = (long long)tiles_x * tiles_y * n;
        const int ctas = (int)(total < 2ll * sms ? total : 2ll * sms);
        corr_fwd_tma<<<ctas, TMA_THREADS, smem, st>>>(tm1, tm2, (float*)out, c, h, w, tiles_x, tiles_y, (int)total);
        return check_launch("correlation_forward(tma)");
      }
    }
    const int smem = 2 * (int)sizeof(CorrStage);
    auto kv = corr_fwd_f32<true>;
    auto ks = corr_fwd_f32<false>;
    auto k = vec ? kv : ks;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("correlation_forward: smem attr: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
    k<<<grid, CORR_THREADS, smem, st>>>((const float*)f1, (const float*)f2, (float*)out, c, h, w);
    return check_launch("correlation_forward");
  }
  corr_fwd<T><<<grid, CORR_THREADS, 0, st>>>((const T*)f1, (const T*)f2, (T*)out, c, h, w);
  return check_launch("correlation_forward");
}

// ---- fp32 backward, tiled -------------------------------------------------------------------------
// gfirst[c,y,x]  = 1/C sum_{dy,dx} gout[k,y,x]       * second[c,y+dy,x+dx]
// gsecond[c,y,x] = 1/C sum_{dy,dx} gout[k,y-dy,x-dx] * first[c,y-dy,x-dx]          k = (dy+4)*9 + (dx+4)
// One kernel for both: a CTA owns an 8x16 pixel tile, stages all 81 gout planes for it ONCE (for gsecond
// each plane pre-shifted by its own displacement, zero outside the image) and then walks the channels 16
// at a time, each with its 16x24 halo tile; a thread keeps 4 pixels x 4 channels in registers, loads the 9
// gout quads of a displacement row once and feeds 144 FMAs from 21 16-byte shared loads.  The first
// version (one thread per gradient element, 81 strided global reads each) ran at 2 % of HBM peak.
constexpr int CB_TH = 8, CB_TW = 16, CB_CK = 16, CB_THREADS = 128;
constexpr int CB_HH = CB_TH + 2 * D, CB_HW = CB_TW + 2 * D;        // 16 x 24 halo
struct CorrBwdSmem {
  float g[ND * ND][CB_TH][CB_TW];       // 41 472 B
  float w[CB_CK][CB_HH][CB_HW];         // 24 576 B
};

template <bool SECOND>
__global__ void __launch_bounds__(CB_THREADS, 3)
corr_bwd_tiled(const float* __restrict__ other, const float* __restrict__ gout, float* __restrict__ gin, int C, int H,
               int W) {
  extern __shared__ __align__(16) uint8_t cb_smem[];
  CorrBwdSmem& S = *reinterpret_cast<CorrBwdSmem*>(cb_smem);
  const int n = blockIdx.z, y0 = blockIdx.y * CB_TH, x0 = blockIdx.x * CB_TW;
  const int tid = threadIdx.x, tq = tid & 3, ty = (tid >> 2) & 7, cs = tid >> 5;
  const size_t plane = (size_t)H * W;
  const float* gn = gout + (size_t)n * (ND * ND) * plane;
  const float* on = other + (size_t)n * C * plane;

  // all 81 gout planes of the tile (SECOND: plane k read at (y - dy, x - dx))
  for (int i = tid; i < ND * ND * CB_TH * CB_TW; i += CB_THREADS) {
    const int xx = i % CB_TW, yy = (i / CB_TW) % CB_TH, k = i / (CB_TW * CB_TH);
    const int dy = k / ND - D, dx = k % ND - D;
    const int gy = y0 + yy - (SECOND ? dy : 0), gx = x0 + xx - (SECOND ? dx : 0);
    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
    cp_async_4_zfill(smem_u32(&S.g[k][yy][xx]), ok ? gn + (size_t)k * plane + (size_t)gy * W + gx : gn, ok);
  }
  cp_async_commit();

  const float inv = 1.f / (float)C;
  const int gy = y0 + ty, gx = x0 + 4 * tq;
  for (int c0 = 0; c0 < C; c0 += CB_CK) {
    __syncthreads();                                  // everyone is done with the previous halo tiles
    for (int i = tid; i < CB_CK * CB_HH * CB_HW; i += CB_THREADS) {
      const int xx = i % CB_HW, yy = (i / CB_HW) % CB_HH, cc = i / (CB_HW * CB_HH);
      const int hy = y0 + yy - D, hx = x0 + xx - D, gc = c0 + cc;
      const bool ok = gc < C && hy >= 0 && hy < H && hx >= 0 && hx < W;
      cp_async_4_zfill(smem_u32(&S.w[cc][yy][xx]), ok ? on + (size_t)gc * plane + (size_t)hy * W + hx : on, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[c][p] = 0.f;
#pragma unroll 1
    for (int dyi = 0; dyi < ND; ++dyi) {
      float g[ND][4];
#pragma unroll
      for (int dxi = 0; dxi < ND; ++dxi) {
        const float4 t = *reinterpret_cast<const float4*>(&S.g[dyi * ND + dxi][ty][4 * tq]);
        g[dxi][0] = t.x; g[dxi][1] = t.y; g[dxi][2] = t.z; g[dxi][3] = t.w;
      }
      const int r = SECOND ? ty + 2 * D - dyi : ty + dyi;     // halo row of the other operand
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float* row = &S.w[cs * 4 + c][r][4 * tq];
        const float4 b0 = *reinterpret_cast<const float4*>(row);
        const float4 b1 = *reinterpret_cast<const float4*>(row + 4);
        const float4 b2 = *reinterpret_cast<const float4*>(row + 8);
        const float b[12] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
        for (int dxi = 0; dxi < ND; ++dxi)
#pragma unroll
          for (int p = 0; p < 4; ++p) acc[c][p] += g[dxi][p] * b[p + (SECOND ? 2 * D - dxi : dxi)];
      }
    }
    if (gy < H && gx < W) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int gc = c0 + cs * 4 + c;
        if (gc < C) {
          float* op = gin + ((size_t)n * C + gc) * plane + (size_t)gy * W + gx;
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (gx + p < W) op[p] = acc[c][p] * inv;
        }
      }
    }
  }
}

template <typename T>
int corr_backward_t(const void* f1, const void* f2, const void* gout, void* g1, void* g2, int n, int c, int h, int w,
                    cudaStream_t st) {
  const long long total = (long long)n * c * h * w;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  int rc = EAVSR_OK;
  if (sizeof(T) == 4 && (long long)h * w > 256 && n <= 65535) {      // tiled fp32 path (not the tiny pyramid maps)
    const int smem = (int)sizeof(CorrBwdSmem);
    dim3 grid(ceil_div(w, CB_TW), ceil_div(h, CB_TH), n);
    if (g1) {
      cudaFuncSetAttribute(corr_bwd_tiled<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      corr_bwd_tiled<false><<<grid, CB_THREADS, smem, st>>>((const float*)f2, (const float*)gout, (float*)g1, c, h, w);
      rc = check_launch("correlation_backward(first, tiled)");
      if (rc) return rc;
    }
    if (g2) {
      cudaFuncSetAttribute(corr_bwd_tiled<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      corr_bwd_tiled<true><<<grid, CB_THREADS, smem, st>>>((const float*)f1, (const float*)gout, (float*)g2, c, h, w);
      rc = check_launch("correlation_backward(second, tiled)");
    }
    return rc;
  }
  if (g1) {
    corr_bwd<T, false><<<blocks, 256, 0, st>>>((const T*)f2, (const T*)gout, (T*)g1, n, c, h, w);
    rc = check_launch("correlation_backward(first)");
    if (rc) return rc;
  }
  if (g2) {
    corr_bwd<T, true><<<blocks, 256, 0, st>>>((const T*)f1, (const T*)gout, (T*)g2, n, c, h, w);
    rc = check_launch("correlation_backward(second)");
  }
  return rc;
}

}  // namespace
}  // namespace eavsr

using namespace eavsr;

extern "C" int eavsr_correlation_forward(const void* first, const void* second, void* out, int n, int c, int h,
                                         int w, int dtype, void* stream) {
  EAVSR_REQUIRE(first && second && out, "correlation_forward: null pointer");
  EAVSR_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "correlation_forward: empty tensor");
  EAVSR_REQUIRE(n <= 65535 && ceil_div(h, TH) <= 65535, "correlation_forward: batch/height too large");
  if (dtype == EAVSR_F32) return corr_forward_t<float>(first, second, out, n, c, h, w, (cudaStream_t)stream);
  if (dtype == EAVSR_BF16) return corr_forward_t<__nv_bfloat16>(first, second, out, n, c, h, w, (cudaStream_t)stream);
  set_error("correlation_forward: bad dtype %d", dtype);
  return EAVSR_ERR_INVALID;
}

extern "C" int eavsr_correlation_backward(const void* first, const void* second, const void* gout, void* gfirst,
                                          void* gsecond, int n, int c, int h, int w, int dtype, void* stream) {
  EAVSR_REQUIRE(first && second && gout, "correlation_backward: null pointer");
  EAVSR_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "correlation_backward: empty tensor");
  if (dtype == EAVSR_F32) return corr_backward_t<float>(first, second, gout, gfirst, gsecond, n, c, h, w, (cudaStream_t)stream);
  if (dtype == EAVSR_BF16) return corr_backward_t<__nv_bfloat16>(first, second, gout, gfirst, gsecond, n, c, h, w, (cudaStream_t)stream);
  set_error("correlation_backward: bad dtype %d", dtype);
  return EAVSR_ERR_INVALID;
}
