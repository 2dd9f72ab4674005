// Generic (any shape / stride / layout) DCNv2 forward and backward for sm_100a.
//
// Replaces mmcv's modulated_deformable_im2col + GEMM (forward) and col2im / col2im_coord /
// im2col + GEMM (backward) that sit under the reference's call at models/networks.py:627-630.
// The column buffer never goes to HBM: each CTA gathers a tile of columns into shared memory
// and contracts it there.  This is the coverage path (arbitrary channels, kernel size, stride,
// dilation, groups, NCHW or NHWC); the model's own configuration (64->64, 3x3, NHWC) runs the
// tcgen05 implicit-GEMM kernel in dcn_fwd_tc.cu for the forward pass.
#include "common.cuh"

namespace eavsr {



namespace {

constexpr int TP = 8;     // output pixels per CTA
constexpr int KC = 512;   // contraction entries staged per pass
constexpr int MAXOB = 8;  // cout/groups <= 32*MAXOB

struct Sample {
  long long o[4];
  float wgt[4];
  float ly, lx;
  bool ok[4];
  bool inside;
};

__device__ __forceinline__ Sample make_sample(float py, float px, int H, int W, const Strides4& xs) {
  Sample s;
  s.inside = (py > -1.f) && (py < (float)H) && (px > -1.f) && (px < (float)W);
  if (!s.inside) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { s.ok[k] = false; s.wgt[k] = 0.f; s.o[k] = 0; }
    s.ly = s.lx = 0.f;
    return s;
  }
  float fy = floorf(py), fx = floorf(px);
  int y0 = (int)fy, x0 = (int)fx, y1 = y0 + 1, x1 = x0 + 1;
  float ly = py - fy, lx = px - fx;
  s.ly = ly; s.lx = lx;
  bool vy0 = y0 >= 0, vy1 = y1 <= H - 1, vx0 = x0 >= 0, vx1 = x1 <= W - 1;
  int cy0 = max(y0, 0), cy1 = min(y1, H - 1), cx0 = max(x0, 0), cx1 = min(x1, W - 1);
  s.ok[0] = vy0 && vx0; s.ok[1] = vy0 && vx1; s.ok[2] = vy1 && vx0; s.ok[3] = vy1 && vx1;
  s.o[0] = cy0 * xs.h + cx0 * xs.w; s.o[1] = cy0 * xs.h + cx1 * xs.w;
  s.o[2] = cy1 * xs.h + cx0 * xs.w; s.o[3] = cy1 * xs.h + cx1 * xs.w;
  s.wgt[0] = s.ok[0] ? (1.f - ly) * (1.f - lx) : 0.f;
  s.wgt[1] = s.ok[1] ? (1.f - ly) * lx : 0.f;
  s.wgt[2] = s.ok[2] ? ly * (1.f - lx) : 0.f;
  s.wgt[3] = s.ok[3] ? ly * lx : 0.f;
  return s;
}

// decode contraction entry kk of conv-group gi at output pixel (n, oy, ox)
struct Entry {
  int c, k;
  size_t off_idx, mask_idx;  // indices into offset (dy; dx = +HO*WO) / mask
  float py0, px0;            // sampling position before the learned offset
};
__device__ __forceinline__ Entry make_entry(const DcnGeom& g, int gi, int kk, int n, int oy, int ox) {
  Entry e;
  const int K = g.KH * g.KW;
  const int cin_g = g.Cin / g.G;
  e.c = gi * cin_g + kk / K;
  e.k = kk % K;
  const int dgi = e.c / (g.Cin / g.DG);
  const size_t P = (size_t)g.HO * g.WO;
  e.off_idx = (((size_t)n * g.DG + dgi) * K + e.k) * 2 * P + (size_t)oy * g.WO + ox;
  e.mask_idx = (((size_t)n * g.DG + dgi) * K + e.k) * P + (size_t)oy * g.WO + ox;
  e.py0 = (float)(oy * g.SH - g.PH + (e.k / g.KW) * g.DH);
  e.px0 = (float)(ox * g.SW - g.PW + (e.k % g.KW) * g.DW);
  return e;
}

template <typename T>
__global__ void __launch_bounds__(256)
dcn_fwd_generic(const T* __restrict__ x, Strides4 xs, const float* __restrict__ offset,
                const float* __restrict__ mask, const T* __restrict__ weight, const T* __restrict__ bias,
                T* __restrict__ out, Strides4 os, DcnGeom g) {
  __shared__ float col[TP][KC];
  const int K = g.KH * g.KW;
  const int cin_g = g.Cin / g.G, cout_g = g.Cout / g.G;
  const int CK = cin_g * K;
  const size_t P = (size_t)g.HO * g.WO;
  const long long total = (long long)g.N * P;
  const long long pix0 = (long long)blockIdx.x * TP;
  const int tp = threadIdx.x >> 5, lane = threadIdx.x & 31;  // phase 2: warp <-> pixel, lane <-> out channel
  for (int gi = 0; gi < g.G; ++gi) {
    float acc[MAXOB];
#pragma unroll
    for (int b = 0; b < MAXOB; ++b) acc[b] = 0.f;
    for (int kc0 = 0; kc0 < CK; kc0 += KC) {
      const int kn = min(KC, CK - kc0);
      __syncthreads();
      for (int idx = threadIdx.x; idx < TP * kn; idx += 256) {
        const int p = idx / kn, kk = kc0 + idx % kn;
        const long long pix = pix0 + p;
        float v = 0.f;
        if (pix < total) {
          const int n = (int)(pix / P);
          const int rem = (int)(pix % P);
          const int oy = rem / g.WO, ox = rem % g.WO;
          Entry e = make_entry(g, gi, kk, n, oy, ox);
          const float dy = __ldg(offset + e.off_idx), dx = __ldg(offset + e.off_idx + P);
          const float m = __ldg(mask + e.mask_idx);
          Sample s = make_sample(e.py0 + dy, e.px0 + dx, g.H, g.W, xs);
          if (s.inside) {
            const T* xc = x + n * xs.n + e.c * xs.c;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (s.ok[q]) v += s.wgt[q] * to_f32<T>(xc[s.o[q]]);
            v *= m;
          }
        }
        col[p][idx % kn] = v;
      }
      __syncthreads();
#pragma unroll
      for (int b = 0; b < MAXOB; ++b) {
        const int ol = b * 32 + lane;
        if (ol < cout_g) {
          const T* wr = weight + ((size_t)(gi * cout_g + ol)) * CK + kc0;
          float a = acc[b];
          for (int kk = 0; kk < kn; ++kk) a += to_f32<T>(wr[kk]) * col[tp][kk];
          acc[b] = a;
        }
      }
    }
    const long long pix = pix0 + tp;
    if (pix < total) {
      const int n = (int)(pix / P);
      const int rem = (int)(pix % P);
      const int oy = rem / g.WO, ox = rem % g.WO;
#pragma unroll
      for (int b = 0; b < MAXOB; ++b) {
        const int ol = b * 32 + lane;
        if (ol < cout_g) {
          const int o = gi * cout_g + ol;
          float r = acc[b] + (bias ? to_f32<T>(bias[o]) : 0.f);
          out[n * os.n + o * os.c + oy * os.h + ox * os.w] = from_f32<T>(r);
        }
      }
    }
  }
}

// Backward wrt x / offset / mask.  Phase A: gcol[p][kk] = sum_o gout[p][o] W[o][kk] into smem;
// phase B: every (p, kk) entry redoes its bilinear gather and scatters its three contributions.
template <typename T>
__global__ void __launch_bounds__(256)
dcn_bwd_data_generic(const T* __restrict__ gout, Strides4 gs, const T* __restrict__ x, Strides4 xs,
                     const float* __restrict__ offset, const float* __restrict__ mask,
                     const T* __restrict__ weight, float* __restrict__ gx32, Strides4 gxs,
                     float* __restrict__ goffset, float* __restrict__ gmask, DcnGeom g) {
  __shared__ float gcol[TP][KC];
  __shared__ float gsm[TP][32 * MAXOB];
  const int K = g.KH * g.KW;
  const int cin_g = g.Cin / g.G, cout_g = g.Cout / g.G;
  const int CK = cin_g * K;
  const size_t P = (size_t)g.HO * g.WO;
  const long long total = (long long)g.N * P;
  const long long pix0 = (long long)blockIdx.x * TP;
  for (int gi = 0; gi < g.G; ++gi) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < TP * cout_g; idx += 256) {
      const int p = idx / cout_g, ol = idx % cout_g;
      const long long pix = pix0 + p;
      float v = 0.f;
      if (pix < total) {
        const int n = (int)(pix / P);
        const int rem = (int)(pix % P);
        v = to_f32<T>(gout[n * gs.n + (gi * cout_g + ol) * gs.c + (rem / g.WO) * gs.h + (rem % g.WO) * gs.w]);
      }
      gsm[p][ol] = v;
    }
    for (int kc0 = 0; kc0 < CK; kc0 += KC) {
      const int kn = min(KC, CK - kc0);
      __syncthreads();
      for (int idx = threadIdx.x; idx < TP * kn; idx += 256) {
        const int p = idx / kn, kl = idx % kn;
        const T* wc = weight + ((size_t)gi * cout_g) * CK + kc0 + kl;
        float a = 0.f;
        for (int ol = 0; ol < cout_g; ++ol) a += gsm[p][ol] * to_f32<T>(wc[(size_t)ol * CK]);
        gcol[p][kl] = a;
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < TP * kn; idx += 256) {
        const int p = idx / kn, kl = idx % kn, kk = kc0 + kl;
        const long long pix = pix0 + p;
        if (pix >= total) continue;
        const int n = (int)(pix / P);
        const int rem = (int)(pix % P);
        const int oy = rem / g.WO, ox = rem % g.WO;
        Entry e = make_entry(g, gi, kk, n, oy, ox);
        const float dy = __ldg(offset + e.off_idx), dx = __ldg(offset + e.off_idx + P);
        const float m = __ldg(mask + e.mask_idx);
        Sample s = make_sample(e.py0 + dy, e.px0 + dx, g.H, g.W, xs);
        if (!s.inside) continue;
        const float gc = gcol[p][kl];
        const T* xc = x + n * xs.n + e.c * xs.c;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = s.ok[q] ? to_f32<T>(xc[s.o[q]]) : 0.f;
        if (gmask) {
          float val = s.wgt[0] * v[0] + s.wgt[1] * v[1] + s.wgt[2] * v[2] + s.wgt[3] * v[3];
          atomicAdd(gmask + e.mask_idx, gc * val);
        }
        const float gm = gc * m;
        if (goffset) {
          float ddy = (v[2] - v[0]) * (1.f - s.lx) + (v[3] - v[1]) * s.lx;
          float ddx = (v[1] - v[0]) * (1.f - s.ly) + (v[3] - v[2]) * s.ly;
          atomicAdd(goffset + e.off_idx, gm * ddy);
          atomicAdd(goffset + e.off_idx + P, gm * ddx);
        }
        if (gx32) {
          float* gxc = gx32 + n * gxs.n + e.c * gxs.c;
          // gx strides may differ from x strides: recompute corner positions
          float fy = floorf(e.py0 + dy), fx = floorf(e.px0 + dx);
          int y0 = (int)fy, x0 = (int)fx;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (s.ok[q]) {
              int yy = y0 + (q >> 1), xx = x0 + (q & 1);
              atomicAdd(gxc + yy * gxs.h + xx * gxs.w, gm * s.wgt[q]);
            }
          }
        }
      }
    }
  }
}

// Backward wrt weight: CTA = 16 contraction entries x all out channels of one conv group,
// looping over pixel chunks (split over gridDim.y); partials land with atomics.
constexpr int KB = 16;
constexpr int PC = 16;
template <typename T>
__global__ void __launch_bounds__(256)
dcn_bwd_weight_generic(const T* __restrict__ gout, Strides4 gs, const T* __restrict__ x, Strides4 xs,
                       const float* __restrict__ offset, const float* __restrict__ mask,
                       float* __restrict__ gweight32, DcnGeom g, int kblocks_per_group) {
  __shared__ float col[PC][KB];
  __shared__ float gsm[PC][32 * MAXOB];
  const int K = g.KH * g.KW;
  const int cin_g = g.Cin / g.G, cout_g = g.Cout / g.G;
  const int CK = cin_g * K;
  const size_t P = (size_t)g.HO * g.WO;
  const long long total = (long long)g.N * P;
  const int gi = blockIdx.x / kblocks_per_group;
  const int kk0 = (blockIdx.x % kblocks_per_group) * KB;
  const int kn = min(KB, CK - kk0);
  constexpr int MAXJ = (32 * MAXOB * KB) / 256;
  float acc[MAXJ];
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) acc[j] = 0.f;
  const long long nchunks = (total + PC - 1) / PC;
  for (long long pc = blockIdx.y; pc < nchunks; pc += gridDim.y) {
    const long long pix0 = pc * PC;
    __syncthreads();
    {
      const int p = threadIdx.x / KB, kl = threadIdx.x % KB;
      const long long pix = pix0 + p;
      float v = 0.f;
      if (pix < total && kl < kn) {
        const int n = (int)(pix / P);
        const int rem = (int)(pix % P);
        Entry e = make_entry(g, gi, kk0 + kl, n, rem / g.WO, rem % g.WO);
        const float dy = __ldg(offset + e.off_idx), dx = __ldg(offset + e.off_idx + P);
        const float m = __ldg(mask + e.mask_idx);
        Sample s = make_sample(e.py0 + dy, e.px0 + dx, g.H, g.W, xs);
        if (s.inside) {
          const T* xc = x + n * xs.n + e.c * xs.c;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (s.ok[q]) v += s.wgt[q] * to_f32<T>(xc[s.o[q]]);
          v *= m;
        }
      }
      col[p][kl] = v;
    }
    for (int idx = threadIdx.x; idx < PC * cout_g; idx += 256) {
      const int p = idx / cout_g, ol = idx % cout_g;
      const long long pix = pix0 + p;
      float v = 0.f;
      if (pix < total) {
        const int n = (int)(pix / P);
        const int rem = (int)(pix % P);
        v = to_f32<T>(gout[n * gs.n + (gi * cout_g + ol) * gs.c + (rem / g.WO) * gs.h + (rem % g.WO) * gs.w]);
      }
      gsm[p][ol] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int idx = threadIdx.x + 256 * j;
      const int ol = idx / KB, kl = idx % KB;
      if (ol < cout_g) {
        float a = acc[j];
#pragma unroll
        for (int p = 0; p < PC; ++p) a += gsm[p][ol] * col[p][kl];
        acc[j] = a;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    const int idx = threadIdx.x + 256 * j;
    const int ol = idx / KB, kl = idx % KB;
    if (ol < cout_g && kl < kn) atomicAdd(gweight32 + ((size_t)(gi * cout_g + ol)) * CK + kk0 + kl, acc[j]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
dcn_bwd_bias(const T* __restrict__ gout, Strides4 gs, float* __restrict__ gbias32, int N, int HO, int WO) {
  __shared__ float red[256];
  const int o = blockIdx.x;
  const long long total = (long long)N * HO * WO;
  float a = 0.f;
  for (long long i = (long long)blockIdx.y * 256 + threadIdx.x; i < total; i += 256ll * gridDim.y) {
    const int n = (int)(i / ((long long)HO * WO));
    const int rem = (int)(i % ((long long)HO * WO));
    a += to_f32<T>(gout[n * gs.n + o * gs.c + (rem / WO) * gs.h + (rem % WO) * gs.w]);
  }
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(gbias32 + o, red[0]);
}

}  // namespace

template <typename T>
int dcn_forward_generic(const void* x, const int64_t* xs, const float* offset, const float* mask,
                        const void* weight, const void* bias, void* out, const int64_t* os, const DcnGeom& g,
                        cudaStream_t st) {
  Strides4 a{xs[0], xs[1], xs[2], xs[3]}, b{os[0], os[1], os[2], os[3]};
  long long total = (long long)g.N * g.HO * g.WO;
  dcn_fwd_generic<T><<<(unsigned)((total + TP - 1) / TP), 256, 0, st>>>((const T*)x, a, offset, mask, (const T*)weight,
                                                                       (const T*)bias, (T*)out, b, g);
  return check_launch("dcn_forward(generic)");
}

template int dcn_forward_generic<float>(const void*, const int64_t*, const float*, const float*, const void*,
                                        const void*, void*, const int64_t*, const DcnGeom&, cudaStream_t);
template int dcn_forward_generic<__nv_bfloat16>(const void*, const int64_t*, const float*, const float*,
                                                const void*, const void*, void*, const int64_t*, const DcnGeom&,
                                                cudaStream_t);

template <typename T>
int dcn_backward_generic(const void* gout, const int64_t* gs, const void* x, const int64_t* xs,
                         const float* offset, const float* mask, const void* weight, float* gx32,
                         const int64_t* gxs, float* goffset, float* gmask, float* gweight32, float* gbias32,
                         const DcnGeom& g, cudaStream_t st) {
  const int K = g.KH * g.KW;
  const size_t P = (size_t)g.HO * g.WO;
  Strides4 sg{gs[0], gs[1], gs[2], gs[3]}, sx{xs[0], xs[1], xs[2], xs[3]}, sgx{0, 0, 0, 0};
  if (gx32) {
    sgx = Strides4{gxs[0], gxs[1], gxs[2], gxs[3]};
    cudaMemsetAsync(gx32, 0, strided_extent_elems(gxs, g.N, g.Cin, g.H, g.W) * sizeof(float), st);
  }
  if (goffset) cudaMemsetAsync(goffset, 0, (size_t)g.N * g.DG * 2 * K * P * sizeof(float), st);
  if (gmask) cudaMemsetAsync(gmask, 0, (size_t)g.N * g.DG * K * P * sizeof(float), st);
  const long long total = (long long)g.N * P;
  int rc = EAVSR_OK;
  if (gx32 || goffset || gmask) {
    dcn_bwd_data_generic<T><<<(unsigned)((total + TP - 1) / TP), 256, 0, st>>>(
        (const T*)gout, sg, (const T*)x, sx, offset, mask, (const T*)weight, gx32, sgx, goffset, gmask, g);
    rc = check_launch("dcn_backward(data)");
    if (rc) return rc;
  }
  if (gweight32) {
    const int CK = (g.Cin / g.G) * K;
    cudaMemsetAsync(gweight32, 0, (size_t)g.Cout * CK * sizeof(float), st);
    const int kb = ceil_div(CK, KB);
    const long long nchunks = (total + PC - 1) / PC;
    int ysplit = (int)(nchunks < 64 ? nchunks : 64);
    dim3 grid(kb * g.G, ysplit);
    dcn_bwd_weight_generic<T><<<grid, 256, 0, st>>>((const T*)gout, sg, (const T*)x, sx, offset, mask, gweight32, g, kb);
    rc = check_launch("dcn_backward(weight)");
    if (rc) return rc;
  }
  if (gbias32) {
    cudaMemsetAsync(gbias32, 0, (size_t)g.Cout * sizeof(float), st);
    int ysplit = (int)((total + 4095) / 4096);
    if (ysplit > 32) ysplit = 32;
    if (ysplit < 1) ysplit = 1;
    dcn_bwd_bias<T><<<dim3(g.Cout, ysplit), 256, 0, st>>>((const T*)gout, sg, gbias32, g.N, g.HO, g.WO);
    rc = check_launch("dcn_backward(bias)");
  }
  return rc;
}

template int dcn_backward_generic<float>(const void*, const int64_t*, const void*, const int64_t*, const float*,
                                         const float*, const void*, float*, const int64_t*, float*, float*, float*,
                                         float*, const DcnGeom&, cudaStream_t);
template int dcn_backward_generic<__nv_bfloat16>(const void*, const int64_t*, const void*, const int64_t*,
                                                 const float*, const float*, const void*, float*, const int64_t*,
                                                 float*, float*, float*, float*, const DcnGeom&, cudaStream_t);

}  // namespace eavsr
