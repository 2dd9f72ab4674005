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

template <typename T>
int corr_forward_t(const void* f1, const void* f2, void* out, int n, int c, int h, int w, cudaStream_t st) {
  dim3 grid(ceil_div(w, TW), ceil_div(h, TH), n);
  corr_fwd<T><<<grid, CORR_THREADS, 0, st>>>((const T*)f1, (const T*)f2, (T*)out, c, h, w);
  return check_launch("correlation_forward");
}

template <typename T>
int corr_backward_t(const void* f1, const void* f2, const void* gout, void* g1, void* g2, int n, int c, int h, int w,
                    cudaStream_t st) {
  const long long total = (long long)n * c * h * w;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  int rc = EAVSR_OK;
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
