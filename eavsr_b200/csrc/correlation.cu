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


// ---- fp32 cost volume on tcgen05 (kind::tf32), straight from the reference's NCHW layout ------------------
// The SIMT kernels above sit at the shared-memory pipe (10 LDS.128 per 108 FMAs and thread, profiles/
// r1_corr_fwd_tma_ncu.txt): the 70 %-of-HBM target needs the multiply-adds on the tensor cores.  Banded GEMM:
// for a 4-row x 32-column block of pixels (M = 128) and the 12 x 32 halo box of `second` that starts 4 pixels up
// and to the left (N = 384),
//     D[pixel m, halo j] = sum_c first[c, m] * second[c, j]                    (K = C, 8 channels per MMA)
// and the 81 outputs of pixel (r, cx) are the entries j = (r + a) * 32 + cx + b, a, b = 0..8.  A 32-wide halo row
// serves pixel columns cx <= 23, so tiles advance by 24 pixels in x (the last 8 MMA rows of every pixel row are
// recomputed by the next tile): ~20 % of D is used, but the tensor pipe has several times the throughput to spare.
//   * Operands are consumed MN-MAJOR, i.e. exactly as NCHW stores them (x contiguous, channel = the K row): one
//     elected thread issues two cp.async.bulk.tensor per 8-channel chunk with tensor maps over (x, c, y, n) --
//     that dimension order makes the TMA unit write [row][channel][32 px] = the canonical MN-major atoms -- zero-
//     filling the image border and the channel tail.  No rearrange / pad pass (the reference runs two,
//     correlation.py:8-33), no transposes.  32-bit MN-major operands exist in ONE swizzle mode only: 128-byte
//     swizzle with 32-byte atomicity (UMMA layout type SWIZZLE_128B_BASE32B, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
//     atoms of 4 K-rows x 128 B, SBO = 512 B between K atoms, LBO = 1 KB between MN atoms); a first version with
//     SWIZZLE_64B / SWIZZLE_128B operands computed nothing at all (the accumulator stayed zero).
//   * fp32 accumulators: 384 TMEM columns.  8 epilogue warps (two per TMEM lane quadrant = pixel row) read a halo
//     row (32 columns) per lane with tcgen05.ld, pick the 9 that belong to their pixel with a 5-level select
//     network on cx (a per-lane register array cannot be indexed dynamically), scale by 1 / C and store -- 24
//     consecutive lanes write 96 contiguous bytes of one (displacement, row).
//   * The band goes through a [81][4][24] shared-memory staging buffer and ONE bulk tensor store per tile (clipped
//     at the image border by the TMA unit), so TMEM is released before the stores drain.
//   * OPT-IN (flag EAVSR_CORR_TF32), not the default, for two measured reasons (30x32x80x128, 178 MB):
//       - kind::tf32 TRUNCATES the inputs to 10 mantissa bits: max-abs error 0.8e-3 (C = 32) ... 2.5e-3 (C = 5) on
//         unit-variance features, i.e. at / over the 1e-3 fp32 bound of BASELINE.json, where the SIMT kernels are exact;
//       - 52 us = the SIMT kernel's time (0.52 of the HBM roofline), not the ~30 us the MMA count promised: loads +
//         MMAs alone take 26 us (the 12 x 32 halo box per 4 x 24 pixels is a 4x amplification of `second`: 230 MB
//         through L2 -> TMA at ~128-byte rows), the band extraction another ~25 us (47-63 selects + 9 stores per
//         halo row and lane), and with 384 of 512 TMEM columns taken by one accumulator the two cannot overlap.
//     What would fix it -- 8-row pixel tiles sharing one halo stage, a second accumulator -- does not fit TMEM.
namespace tc {
constexpr int PT_H = 4, PT_W = 32, PT_STEP = 24;   // MMA rows per tile: 4 x 32 pixels, of which 4 x 24 are stored
constexpr int HB_H = PT_H + 2 * D, HB_W = 32;       // halo box (12 x 32)
constexpr int KC = 8;                              // channels per stage = one kind::tf32 K step
constexpr int A_BYTES = PT_H * KC * PT_W * 4;      // 4 KB   [y][c][32 px]
constexpr int B_BYTES = HB_H * KC * HB_W * 4;      // 12 KB  [hy][c][32 px]
constexpr int STAGE = A_BYTES + B_BYTES;           // 16 KB
constexpr int NST = 8;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (2 + EPI_WARPS) * 32;      // warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue
constexpr int NCOLS = HB_H * HB_W;                 // 384 accumulator columns
constexpr int OUT_BYTES = ND * ND * PT_H * PT_STEP * 4;   // 31 104 B: [81][4 rows][24 px] of one tile, TMA-stored
constexpr int OUT_OFF = NST * STAGE;               // two output staging buffers (128-byte aligned)
constexpr int OUT_PITCH = (OUT_BYTES + 127) / 128 * 128;
constexpr int BAR_OFF = OUT_OFF + 2 * OUT_PITCH;
constexpr int BAR_BYTES = (2 * NST + 2) * 8 + 16;
constexpr int DYN = BAR_OFF + BAR_BYTES + 1024;
static_assert(DYN <= 232448, "shared memory budget");

// 32-bit MN-major operand, SWIZZLE_128B_BASE32B: atoms of 4 K-rows x 128 B (32 elements along MN);
// LBO = byte stride between atoms along MN, SBO = byte stride between atoms along K
__device__ __forceinline__ uint64_t desc_mn32(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                          // LayoutType::SWIZZLE_128B_BASE32B
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32_mn(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

#ifdef EAVSR_CORR_DEBUG     // development only: stage 0 as the TMA unit wrote it + the accumulator of CTA 0's first tile
__device__ float g_corr_dbg[STAGE / 4 + 128 * NCOLS];
#endif

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int x, int y, int d, int n) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(src), "r"(x), "r"(y), "r"(d), "r"(n)
               : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
corr_fwd_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmO, int C, int H, int W, int tiles_x, int tiles_y, int total_tiles) {
  extern __shared__ uint8_t tc_smem[];
  const uint32_t sbase = (smem_u32(tc_smem) + 1023u) & ~1023u;
  uint8_t* sgen = tc_smem + (sbase - smem_u32(tc_smem));
  const uint32_t bars = sbase + BAR_OFF;
  const uint32_t bar_full = bars, bar_empty = bars + NST * 8, bar_accf = bars + 2 * NST * 8, bar_acce = bar_accf + 8;
  const uint32_t tmem_slot_addr = bar_acce + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + BAR_OFF + (2 * NST + 2) * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_accf, 1);
    mbar_init(bar_acce, EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const int nchunks = (C + KC - 1) / KC;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
#define TILE_OF(tl) ((int)blockIdx.x + (tl) * (int)gridDim.x)
  const int tiles_per_img = tiles_x * tiles_y;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int cnt = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int tile = TILE_OF(tl);
        const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
        const int y0 = (rem / tiles_x) * PT_H, x0 = (rem % tiles_x) * PT_STEP;
        for (int k = 0; k < nchunks; ++k, ++cnt) {
          const int s = cnt % NST;
          if (cnt >= NST) mbar_wait(bar_empty + 8 * s, ((cnt / NST) - 1) & 1);
          const uint32_t full = bar_full + 8 * s, dst = sbase + s * STAGE;
          mbar_arrive_expect_tx(full, STAGE);
          // tensor-map dimensions are (x, c, y, n)
          tma_load_4d(dst, &tmB, x0 - D, k * KC, y0 - D, n, full);
          tma_load_4d(dst + B_BYTES, &tmA, x0, k * KC, y0, n, full);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t IDESC_256 = idesc_tf32_mn(128, 256), IDESC_128 = idesc_tf32_mn(128, 128);
      int cnt = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        if (tl >= 1) mbar_wait(bar_acce, (tl - 1) & 1);        // the epilogue has read tile tl-1 out of TMEM
        tc_fence_after();
        for (int k = 0; k < nchunks; ++k, ++cnt) {
          const int s = cnt % NST;
          mbar_wait(bar_full + 8 * s, (cnt / NST) & 1);
          tc_fence_after();
          const uint32_t sb = sbase + s * STAGE, sa = sb + B_BYTES;
          const uint64_t adesc = desc_mn32(sa, 1024, 512);
          const uint64_t bdesc = desc_mn32(sb, 1024, 512);
          umma_tf32(tmem_d, adesc, bdesc, IDESC_256, k != 0);                                    // halo rows 0-7
          umma_tf32(tmem_d + 256, adesc, bdesc + (uint64_t)((8 * 1024) >> 4), IDESC_128, k != 0);   // halo rows 8-11
          umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_accf);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: band extraction -> shared memory -> TMA store =====================
    // Stores straight from registers (81 planes x 96-byte pieces per tile) cost more than everything else together
    // (ablation: 72 us with them, 29 us without), and with a single 384-column accumulator they also held up the
    // next tile's MMAs.  The band goes to a [81][4][24] staging buffer instead, TMEM is released, and ONE bulk
    // tensor store per tile (clipped at the image border by the TMA unit) drains it while the next tile runs.
    const int q = warp & 3;                         // TMEM lane quadrant = pixel row of the tile
    const int half = (warp - 2) >> 2;               // halo rows q .. q+4 (half 0) / q+5 .. q+8 (half 1)
    const int cx = lane;                            // pixel column (kept when < PT_STEP)
    const float inv = 1.f / (float)C;
    const bool b4 = cx & 16, b3 = cx & 8, b2 = cx & 4, b1 = cx & 2, b0 = cx & 1;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int tile = TILE_OF(tl);
      const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
      const int y0 = (rem / tiles_x) * PT_H, x0 = (rem % tiles_x) * PT_STEP;
      float* so = reinterpret_cast<float*>(sgen + OUT_OFF + (tl & 1) * OUT_PITCH) + q * PT_STEP + cx;
      mbar_wait(bar_accf, tl & 1);
      tc_fence_after();
#ifdef EAVSR_CORR_DEBUG
      if (blockIdx.x == 0 && tl == 0) {
        if (half == 0) {
          for (int col = 0; col < NCOLS; col += 32) {
            uint32_t t32[32];
            tmem_ld_32x32(tmem_d + ((uint32_t)(q * 32) << 16) + col, t32);
            tmem_ld_wait();
            for (int e = 0; e < 32; ++e) g_corr_dbg[STAGE / 4 + (q * 32 + lane) * NCOLS + col + e] = __uint_as_float(t32[e]);
          }
        }
        if (warp == 2)
          for (int e = lane; e < STAGE / 4; e += 32) g_corr_dbg[e] = reinterpret_cast<const float*>(sgen)[e];
      }
#endif
      const int a_lo = half * 5, a_hi = half ? ND : 5;
#pragma unroll 1
      for (int a = a_lo; a < a_hi; ++a) {           // displacement row a <-> halo row q + a (warp-uniform)
        uint32_t v[32];
        tmem_ld_32x32(tmem_d + ((uint32_t)(q * 32) << 16) + (q + a) * HB_W, v);
        tmem_ld_wait();
        // w0[b] = v[cx + b], cx = 0..23: shift by 16 (only cx >= 16, i.e. then by < 8 more), 8, 4, 2, 1
        uint32_t w4[24], w3[16], w2[12], w1[10], w0[9];
#pragma unroll
        for (int j = 0; j < 16; ++j) w4[j] = b4 ? v[j + 16] : v[j];
#pragma unroll
        for (int j = 16; j < 24; ++j) w4[j] = v[j];
#pragma unroll
        for (int j = 0; j < 16; ++j) w3[j] = b3 ? w4[j + 8] : w4[j];
#pragma unroll
        for (int j = 0; j < 12; ++j) w2[j] = b2 ? w3[j + 4] : w3[j];
#pragma unroll
        for (int j = 0; j < 10; ++j) w1[j] = b1 ? w2[j + 2] : w2[j];
#pragma unroll
        for (int j = 0; j < 9; ++j) w0[j] = b0 ? w1[j + 1] : w1[j];
        if (cx < PT_STEP) {
#pragma unroll
          for (int b = 0; b < ND; ++b) so[(a * ND + b) * (PT_H * PT_STEP)] = __uint_as_float(w0[b]) * inv;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce);          // TMEM is free: the next tile's MMAs may start
      fence_proxy_async_smem();                      // staging writes -> visible to the TMA store
      if (warp == 2 && lane == 0)                    // the store of tile tl-1 has read its buffer: tile tl+1 may fill it
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      asm volatile("bar.sync 1, %0;\n" ::"n"(EPI_WARPS * 32) : "memory");
      if (warp == 2 && lane == 0) {
        tma_store_4d(&tmO, sbase + OUT_OFF + (tl & 1) * OUT_PITCH, x0, y0, 0, n);
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
    }
    if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_d);
}

// (x, c, y, n) view of an NCHW fp32 tensor with a [32 x KC x box_h x 1] box: the TMA unit then writes
// [row][channel][32 px], the MN-major operand layout; zero fill outside (image border, channel tail)
bool make_map(CUtensorMap* tm, const void* base, int n, int c, int h, int w, int box_h) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)w, (cuuint64_t)c, (cuuint64_t)h, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)w * h * 4, (cuuint64_t)w * 4, (cuuint64_t)w * h * c * 4};
  const cuuint32_t box[4] = {32, (cuuint32_t)KC, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns -1 when this shape is not eligible (caller falls back to the SIMT kernels)
int launch(const void* f1, const void* f2, void* out, int n, int c, int h, int w, cudaStream_t st) {
  const bool ok = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(f1) & 15u) == 0) &&
                  ((reinterpret_cast<uintptr_t>(f2) & 15u) == 0) && (long long)w * h * c * 4 < (1ll << 40) &&
                  (long long)h * w > 256 && (long long)n * 81 * h * w < (1ll << 40);
  if (!ok) return -1;
  if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0) return -1;
  CUtensorMap tmA, tmB, tmO;
  if (!make_map(&tmA, f1, n, c, h, w, PT_H) || !make_map(&tmB, f2, n, c, h, w, HB_H)) return -1;
  {  // (x, y, displacement, n) view of the output with a [24 x 4 x 81 x 1] box, no swizzle
    EncodeTiledFn enc = encode_tiled_fn();
    const cuuint64_t dims[4] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)(ND * ND), (cuuint64_t)n};
    const cuuint64_t strides[3] = {(cuuint64_t)w * 4, (cuuint64_t)w * h * 4, (cuuint64_t)w * h * ND * ND * 4};
    const cuuint32_t box[4] = {(cuuint32_t)PT_STEP, (cuuint32_t)PT_H, (cuuint32_t)(ND * ND), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (!enc || enc(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) !=
                    CUDA_SUCCESS)
      return -1;
  }
  cudaError_t e = cudaFuncSetAttribute(corr_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN);
  if (e != cudaSuccess) { set_error("correlation_forward(tc): smem attr: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles_x = ceil_div(w, PT_STEP), tiles_y = ceil_div(h, PT_H);
  const long long total = (long long)tiles_x * tiles_y * n;
  if (total >= (1ll << 31)) return -1;
  const int ctas = (int)(total < sms ? total : sms);
  corr_fwd_tc<<<ctas, THREADS, DYN, st>>>(tmA, tmB, tmO, c, h, w, tiles_x, tiles_y, (int)total);
  return check_launch("correlation_forward(tcgen05)");
}
}  // namespace tc

template <typename T>
int corr_forward_t(const void* f1, const void* f2, void* out, int n, int c, int h, int w, cudaStream_t st, unsigned flags = 0) {
  if (h * w <= 256) {                       // PWC pyramid of 64x64 training crops: 1x1 ... 16x16
    const long long total = (long long)n * ND * ND * h * w;
    corr_fwd_small<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const T*)f1, (const T*)f2, (T*)out, n, c, h, w);
    return check_launch("correlation_forward(small)");
  }
  dim3 grid(ceil_div(w, TW), ceil_div(h, TH), n);
  if (sizeof(T) == 4 && (flags & EAVSR_CORR_TF32)) {       // opt-in: tcgen05 / tf32 (see the kernel's header comment)
    const int rc = tc::launch(f1, f2, out, n, c, h, w, st);
    if (rc >= 0) return rc;
  }
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

extern "C" int eavsr_correlation_forward_ex(const void* first, const void* second, void* out, int n, int c, int h,
                                            int w, int dtype, unsigned flags, void* stream) {
  EAVSR_REQUIRE(first && second && out, "correlation_forward: null pointer");
  EAVSR_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "correlation_forward: empty tensor");
  EAVSR_REQUIRE(n <= 65535 && ceil_div(h, TH) <= 65535, "correlation_forward: batch/height too large");
  if (dtype == EAVSR_F32) return corr_forward_t<float>(first, second, out, n, c, h, w, (cudaStream_t)stream, flags);
  if (dtype == EAVSR_BF16) return corr_forward_t<__nv_bfloat16>(first, second, out, n, c, h, w, (cudaStream_t)stream, flags);
  set_error("correlation_forward: bad dtype %d", dtype);
  return EAVSR_ERR_INVALID;
}

extern "C" int eavsr_correlation_forward(const void* first, const void* second, void* out, int n, int c, int h,
                                         int w, int dtype, void* stream) {
  return eavsr_correlation_forward_ex(first, second, out, n, c, h, w, dtype, 0u, stream);
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

#ifdef EAVSR_CORR_DEBUG
extern "C" int eavsr_debug_corr(float* host, int count) {
  return (int)cudaMemcpyFromSymbol(host, eavsr::tc::g_corr_dbg, sizeof(float) * count);
}
#endif
