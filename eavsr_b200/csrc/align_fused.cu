// Fused producers / consumers around the alignment hot path (inference, sm_100a).
//
// ncu launch list of the full EAVSR+ forward (profiles/): once DCNv2 and the warps run on this
// library, 75 % of the device time is PyTorch glue between them -- cuDNN's grouped-direct kernel for
// the two tiny grouped 3x3 convolutions in every AdaptBlock (403 us per call), TensorIterator
// broadcast kernels for the affine offset expansion, and a reduce + four small launches per channel
// attention.  These kernels replace that glue:
//
//   adapt_mix        concat2(concat(cat[x, ref]))            models/networks.py:289-290,299 / 327-328,334
//                    = depthwise 3x3 (128 ch) + LeakyReLU(0.2) + grouped 3x3 (2 -> 1, 64 groups) +
//                    LeakyReLU(0.2) in ONE pass over the two inputs (50 MB instead of ~200 MB + 2 launches)
//   affine_offsets   T*R - R + t expansion + sigmoid(mask)   models/networks.py:302-313
//   channel_sum /    CALayer (global mean -> 1x1 -> ReLU -> 1x1 -> sigmoid -> scale) + residual add
//   ca_scale_residual                                        models/networks.py:449-465 (RCABlock)
#include "common.cuh"

namespace eavsr {
namespace {

// ---------------------------------------------------------------------------------------------
// adapt_mix
// ---------------------------------------------------------------------------------------------
constexpr int MX_TH = 8, MX_TW = 16;            // output tile
constexpr int MX_IH = MX_TH + 4, MX_IW = MX_TW + 4;
constexpr int MX_YH = MX_TH + 2, MX_YW = MX_TW + 2;
constexpr int MX_C = 64;                        // channels of each input
constexpr int MX_THREADS = 256;

// 8 consecutive channels from shared memory, widened to fp32 (one / two 16-byte LDS)
__device__ __forceinline__ void lds8(const __nv_bfloat16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  f[0] = bf16lo_to_f32(u.x); f[1] = bf16hi_to_f32(u.x); f[2] = bf16lo_to_f32(u.y); f[3] = bf16hi_to_f32(u.y);
  f[4] = bf16lo_to_f32(u.z); f[5] = bf16hi_to_f32(u.z); f[6] = bf16lo_to_f32(u.w); f[7] = bf16hi_to_f32(u.w);
}
__device__ __forceinline__ void lds8(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

template <typename T> struct MixSmem {
  T in[MX_IH * MX_IW][MX_C];
  T y[MX_YH * MX_YW][MX_C];
  float w1[9][MX_C];
  float w2[9][MX_C];
  float b1[MX_C];
  float b2[MX_C / 2];
};

template <typename T>
__global__ void __launch_bounds__(MX_THREADS)
adapt_mix_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ w1g,
                 const T* __restrict__ b1g, const T* __restrict__ w2g, const T* __restrict__ b2g,
                 T* __restrict__ out, int H, int W, float slope) {
  extern __shared__ __align__(16) uint8_t mix_smem_raw[];
  MixSmem<T>& S = *reinterpret_cast<MixSmem<T>*>(mix_smem_raw);
  constexpr int VEC = 16 / sizeof(T);           // channels per 16-byte chunk
  constexpr int CPP = MX_C / VEC;               // chunks per pixel
  const int half = blockIdx.z & 1, n = blockIdx.z >> 1;
  const int y0 = blockIdx.y * MX_TH, x0 = blockIdx.x * MX_TW;
  const T* src = (half ? b : a) + (size_t)n * H * W * MX_C;
  const int tid = threadIdx.x;

  // weights of this half: cat channels [half*64, half*64+64) -> out channels [half*32, half*32+32)
  for (int i = tid; i < 9 * MX_C; i += MX_THREADS) {
    const int tap = i / MX_C, c = i % MX_C;
    S.w1[tap][c] = to_f32<T>(w1g[(size_t)(half * MX_C + c) * 9 + tap]);
    // y channel c feeds out channel half*32 + c/2 as its input (c & 1)
    S.w2[tap][c] = to_f32<T>(w2g[((size_t)(half * (MX_C / 2) + (c >> 1)) * 2 + (c & 1)) * 9 + tap]);
  }
  for (int i = tid; i < MX_C; i += MX_THREADS) S.b1[i] = to_f32<T>(b1g[half * MX_C + i]);
  for (int i = tid; i < MX_C / 2; i += MX_THREADS) S.b2[i] = to_f32<T>(b2g[half * (MX_C / 2) + i]);

  // input tile with a 2-pixel halo, zero outside the image (conv zero padding)
  // (cp.async: all 7-8 copies of a thread in flight at once -- with a load -> store loop the tile fill was a
  // chain of dependent L2 round trips and 28 % of the kernel's stall samples sat on its STS)
  for (int i = tid; i < MX_IH * MX_IW * CPP; i += MX_THREADS) {
    const int p = i / CPP, ch = i % CPP;
    const int gy = y0 - 2 + p / MX_IW, gx = x0 - 2 + p % MX_IW;
    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
    const T* g = ok ? src + ((size_t)gy * W + gx) * MX_C + ch * VEC : src;
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(&S.in[p][ch * VEC])), "l"(g), "r"(sz));
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  // Every thread works on ONE 8-channel chunk for the whole kernel (256 % 8 == 0), so its 9x8 weights
  // live in registers: the inner loops are one 16-byte LDS + 8 FMAs per tap.
  const int c0 = (tid & 7) * 8;
  float wr[9][8], br[8];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int e = 0; e < 8; ++e) wr[tap][e] = S.w1[tap][c0 + e];
#pragma unroll
  for (int e = 0; e < 8; ++e) br[e] = S.b1[c0 + e];

  // Both phases are instruction bound (bf16 -> fp32 widening is 40 % of the work), so every thread computes
  // TWO horizontally adjacent pixels per item: 3 x 4 chunk loads and conversions feed 2 x 72 FMAs instead
  // of 2 x (3 x 3), and results leave as one 16-byte / 8-byte store per pixel.
  auto st8 = [](T* p, const float* f) {
    if constexpr (sizeof(T) == 2) {
      VecLoad<T, 8>::st(p, f);
    } else {
      *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
  };
  // phase 1: y = lrelu(depthwise3x3(in) + b1) on the (TH+2)x(TW+2) ring, 0 outside the image
  static_assert(MX_YW % 2 == 0 && MX_TW % 2 == 0, "pixel pairs");
  for (int i = tid; i < MX_YH * (MX_YW / 2) * (MX_C / 8); i += MX_THREADS) {
    const int pp = i / (MX_C / 8);
    const int yy = pp / (MX_YW / 2), xx = (pp % (MX_YW / 2)) * 2;
    float acc[2][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[0][e] = acc[1][e] = br[e];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float v[4][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) lds8(&S.in[(yy + r) * MX_IW + xx + c][c0], v[c]);
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          acc[0][e] += wr[r * 3 + t][e] * v[t][e];
          acc[1][e] += wr[r * 3 + t][e] * v[t + 1][e];
        }
    }
    const int gy = y0 - 1 + yy;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int gx = x0 - 1 + xx + k;
      const bool inside = gy >= 0 && gy < H && gx >= 0 && gx < W;
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float v = acc[k][e] > 0.f ? acc[k][e] : acc[k][e] * slope;
        o[e] = inside ? v : 0.f;
      }
      st8(&S.y[yy * MX_YW + xx + k][c0], o);
    }
  }
  __syncthreads();

  // phase 2: out[o] = lrelu(b2[o] + sum_tap w2[tap][2o] y[2o] + w2[tap][2o+1] y[2o+1])
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int e = 0; e < 8; ++e) wr[tap][e] = S.w2[tap][c0 + e];
  for (int i = tid; i < MX_TH * (MX_TW / 2) * (MX_C / 8); i += MX_THREADS) {
    const int pp = i / (MX_C / 8);
    const int yy = pp / (MX_TW / 2), xx = (pp % (MX_TW / 2)) * 2;
    const int gy = y0 + yy, gx = x0 + xx;
    if (gy >= H || gx >= W) continue;
    float acc[2][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[0][e] = acc[1][e] = S.b2[(c0 >> 1) + e];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float v[4][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) lds8(&S.y[(yy + r) * MX_YW + xx + c][c0], v[c]);
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          acc[0][e >> 1] += wr[r * 3 + t][e] * v[t][e];
          acc[1][e >> 1] += wr[r * 3 + t][e] * v[t + 1][e];
        }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (gx + k >= W) continue;
      T* op = out + ((size_t)n * H * W + (size_t)gy * W + gx + k) * MX_C + half * (MX_C / 2) + (c0 >> 1);
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = acc[k][e] > 0.f ? acc[k][e] : acc[k][e] * slope;
      if constexpr (sizeof(T) == 2) {
        *reinterpret_cast<uint2*>(op) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
      } else {
        *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// affine offsets + sigmoid mask.  One thread per (pixel, group); outputs are fp32 NCHW planes
// (lanes along pixels -> coalesced), exactly the layout eavsr_dcn_forward consumes.
// offset[(g*9+k)*2+i] = T[g][i][0]*R0[k] + T[g][i][1]*R1[k] - R_i[k] + t[g][i]
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
affine_offsets_kernel(const T* __restrict__ tm, Strides4 ts, const T* __restrict__ tr, Strides4 rs,
                      const T* __restrict__ ml, Strides4 ms, const T* __restrict__ tm_bias,
                      const T* __restrict__ tr_bias, const T* __restrict__ ml_bias, float* __restrict__ offset,
                      float* __restrict__ mask, int N, int D, int H, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long HW = (long long)H * W;
  if (idx >= (long long)N * D * HW) return;
  const int pix = (int)(idx % HW);
  const int g = (int)((idx / HW) % D);
  const int n = (int)(idx / (HW * D));
  const int y = pix / W, x = pix % W;
  const T* tp = tm + n * ts.n + y * ts.h + x * ts.w + (long long)(g * 4) * ts.c;
  const T* rp = tr + n * rs.n + y * rs.h + x * rs.w + (long long)(g * 2) * rs.c;
  float a = to_f32<T>(tp[0]), b = to_f32<T>(tp[ts.c]), c = to_f32<T>(tp[2 * ts.c]), d = to_f32<T>(tp[3 * ts.c]);
  float t0 = to_f32<T>(rp[0]), t1 = to_f32<T>(rp[rs.c]);
  if (tm_bias) {   // biases of the bias-free convolutions that produced T / t / logits
    a += to_f32<T>(tm_bias[g * 4]); b += to_f32<T>(tm_bias[g * 4 + 1]);
    c += to_f32<T>(tm_bias[g * 4 + 2]); d += to_f32<T>(tm_bias[g * 4 + 3]);
  }
  if (tr_bias) { t0 += to_f32<T>(tr_bias[g * 2]); t1 += to_f32<T>(tr_bias[g * 2 + 1]); }
  float* op = offset + ((size_t)(n * D + g) * 18) * HW + pix;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float r0 = (float)(k / 3 - 1), r1 = (float)(k % 3 - 1);
    op[(size_t)(2 * k) * HW] = a * r0 + b * r1 - r0 + t0;
    op[(size_t)(2 * k + 1) * HW] = c * r0 + d * r1 - r1 + t1;
  }
  if (mask) {
    const T* mp = ml + n * ms.n + y * ms.h + x * ms.w + (long long)(g * 9) * ms.c;
    float* mo = mask + ((size_t)(n * D + g) * 9) * HW + pix;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float lg = to_f32<T>(mp[k * ms.c]) + (ml_bias ? to_f32<T>(ml_bias[g * 9 + k]) : 0.f);
      mo[(size_t)k * HW] = 1.f / (1.f + __expf(-lg));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// channel attention: sums over H*W (NHWC), then out = res * sigmoid(MLP(mean)) + skip
// ---------------------------------------------------------------------------------------------
constexpr int CS_THREADS = 256;
constexpr int CS_PIX = 512;   // pixels per CTA

template <typename T, int C, bool PROD = false>
__global__ void __launch_bounds__(CS_THREADS)
channel_sum_kernel(const T* __restrict__ x, float* __restrict__ sums, int HW, const T* __restrict__ x2 = nullptr) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int CPP = C / VEC;
  constexpr int PPI = CS_THREADS / CPP;          // pixels per iteration
  __shared__ float red[PPI][C];
  const int n = blockIdx.y;
  const int ch = threadIdx.x % CPP, pl = threadIdx.x / CPP;
  const int p0 = blockIdx.x * CS_PIX;
  const int p1 = min(p0 + CS_PIX, HW);
  float acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
  const T* xn = x + (size_t)n * HW * C;
  for (int p = p0 + pl; p < p1; p += PPI) {
    float v[VEC];
    VecLoad<T, VEC>::ld(xn + (size_t)p * C + ch * VEC, v);
    if (PROD) {                                    // sums of the element-wise product of two tensors
      float u[VEC];
      VecLoad<T, VEC>::ld(x2 + (size_t)n * HW * C + (size_t)p * C + ch * VEC, u);
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[e] += v[e] * u[e];
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[e] += v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) red[pl][ch * VEC + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < C) {
    float s = 0.f;
#pragma unroll 4
    for (int i = 0; i < PPI; ++i) s += red[i][threadIdx.x];
    atomicAdd(sums + n * C + threadIdx.x, s);
  }
}

template <typename T, int C, int R>
__global__ void __launch_bounds__(256)
ca_scale_residual_kernel(const T* __restrict__ res, const T* __restrict__ skip, const float* __restrict__ sums,
                         const T* __restrict__ w1, const T* __restrict__ b1, const T* __restrict__ w2,
                         const T* __restrict__ b2, const T* __restrict__ res_bias, T* __restrict__ out, int HW,
                         float inv_hw, int chunks_per_img) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int CPP = C / VEC;
  __shared__ float part[8];
  __shared__ float hid[R];
  __shared__ float scale[C];
  __shared__ float rb[C];                       // bias of the conv that produced `res` (0 if none)
  static_assert(R * C == 256, "one thread per (hidden unit, channel) product");
  const int n = blockIdx.y;
  // The squeeze-excite MLP is a prologue every CTA pays before it can stream: all 256 products of the
  // first layer are loaded and multiplied in parallel (one global round trip) and reduced with shuffles,
  // instead of four threads walking 64 dependent loads each (~3 us of a 12 us kernel).
  {
    const int r = threadIdx.x >> 6, i = threadIdx.x & 63;
    const float bi = res_bias ? to_f32<T>(res_bias[i]) : 0.f;
    if (r == 0) rb[i] = bi;
    float p = to_f32<T>(w1[r * C + i]) * (sums[n * C + i] * inv_hw + bi);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = p;
  }
  __syncthreads();
  if (threadIdx.x < R) hid[threadIdx.x] = fmaxf(to_f32<T>(b1[threadIdx.x]) + part[2 * threadIdx.x] + part[2 * threadIdx.x + 1], 0.f);
  __syncthreads();
  if (threadIdx.x < C) {
    float a = to_f32<T>(b2[threadIdx.x]);
#pragma unroll
    for (int j = 0; j < R; ++j) a += to_f32<T>(w2[threadIdx.x * R + j]) * hid[j];
    scale[threadIdx.x] = 1.f / (1.f + __expf(-a));
  }
  __syncthreads();
  const size_t base = (size_t)n * HW * C;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < chunks_per_img; i += gridDim.x * 256) {
    const int ch = i % CPP;
    float r[VEC], s[VEC], o[VEC];
    VecLoad<T, VEC>::ld(res + base + (size_t)i * VEC, r);
    VecLoad<T, VEC>::ld(skip + base + (size_t)i * VEC, s);
#pragma unroll
    for (int e = 0; e < VEC; ++e) o[e] = (r[e] + rb[ch * VEC + e]) * scale[ch * VEC + e] + s[e];
    VecLoad<T, VEC>::st(out + base + (size_t)i * VEC, o);
  }
}

// x[p][c] = act(x[p][c] + bias[c]) in place on a dense NHWC tensor; act = LeakyReLU(slope)
// (slope 1 = identity, 0 = ReLU).  PyTorch adds a cuDNN convolution's bias with a broadcast
// TensorIterator kernel (not vectorised for channels_last: 17.6 us on a 64x272x480 bf16 map) and the
// activation with another launch; this is one 16-byte-vectorised pass.
template <typename T>
__global__ void __launch_bounds__(256)
bias_act_kernel(T* __restrict__ x, const T* __restrict__ bias, int C, long long nchunks, float slope) {
  constexpr int VEC = 16 / sizeof(T);
  const int cpp = C / VEC;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nchunks; i += (long long)gridDim.x * 256) {
    const int ch = (int)(i % cpp);
    float v[VEC], b[VEC];
    VecLoad<T, VEC>::ld(x + i * VEC, v);
    VecLoad<T, VEC>::ld(bias + ch * VEC, b);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float t = v[e] + b[e];
      v[e] = t > 0.f ? t : t * slope;
    }
    VecLoad<T, VEC>::st(x + i * VEC, v);
  }
}

// PixelShuffle(2) of a dense NHWC tensor with the producing convolution's bias and LeakyReLU folded
// in: out[n, 2h+i, 2w+j, c] = act(x[n, h, w, 4c + 2i + j] + bias[4c + 2i + j]).  One thread reads the
// 4*VEC consecutive input channels that make VEC output channels of the four output pixels (64 B)
// and writes four 16-byte vectors.  Replaces bias add + activation + pixel_shuffle + two layout copies
// (1.2 ms per 544x960 -> 1088x1920 frame in the ATen path).
template <typename T>
__global__ void __launch_bounds__(256)
bias_act_shuffle_kernel(const T* __restrict__ x, const T* __restrict__ bias, T* __restrict__ out, int Cout, int H,
                        int W, long long items, float slope) {
  constexpr int VEC = 16 / sizeof(T);
  const int cpp = Cout / VEC;                     // output chunks per input pixel
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < items; i += (long long)gridDim.x * 256) {
    const int ch = (int)(i % cpp);
    const long long pix = i / cpp;
    const int w = (int)(pix % W);
    const long long nh = pix / W;                 // n * H + h
    const T* ip = x + pix * (4 * Cout) + ch * (4 * VEC);
    float v[4 * VEC], b[4 * VEC];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      VecLoad<T, VEC>::ld(ip + k * VEC, v + k * VEC);
      if (bias) VecLoad<T, VEC>::ld(bias + ch * (4 * VEC) + k * VEC, b + k * VEC);
    }
#pragma unroll
    for (int e = 0; e < 4 * VEC; ++e) {
      const float t = v[e] + (bias ? b[e] : 0.f);
      v[e] = t > 0.f ? t : t * slope;
    }
#pragma unroll
    for (int ij = 0; ij < 4; ++ij) {
      float o[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] = v[e * 4 + ij];
      const long long orow = nh * 2 + (ij >> 1);  // n * 2H + 2h + i
      T* op = out + ((orow * (2 * W)) + 2 * w + (ij & 1)) * Cout + ch * VEC;
      VecLoad<T, VEC>::st(op, o);
    }
  }
}

bool dense_nhwc(const int64_t s[4], int c, int h, int w) {
  return s[1] == 1 && s[3] == c && s[2] == (int64_t)w * c && s[0] == (int64_t)h * w * c;
}
bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
int mix_launch(const void* a, const void* b, const void* w1, const void* b1, const void* w2, const void* b2,
               void* out, int n, int h, int w, float slope, cudaStream_t st) {
  auto k = adapt_mix_kernel<T>;
  const int smem = (int)sizeof(MixSmem<T>);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("adapt_mix: smem attr: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
  dim3 grid(ceil_div(w, MX_TW), ceil_div(h, MX_TH), n * 2);
  k<<<grid, MX_THREADS, smem, st>>>((const T*)a, (const T*)b, (const T*)w1, (const T*)b1, (const T*)w2,
                                    (const T*)b2, (T*)out, h, w, slope);
  return check_launch("adapt_mix");
}

}  // namespace
}  // namespace eavsr

using namespace eavsr;

extern "C" int eavsr_adapt_mix_forward(const void* a, const void* b, const void* w1, const void* b1,
                                       const void* w2, const void* b2, void* out, int n, int c, int h, int w,
                                       float negative_slope, int dtype, void* stream) {
  EAVSR_REQUIRE(a && b && w1 && b1 && w2 && b2 && out, "adapt_mix: null pointer");
  EAVSR_REQUIRE(n > 0 && h > 0 && w > 0, "adapt_mix: empty tensor");
  if (c != MX_C) { set_error("adapt_mix: only %d-channel inputs are fused (got %d)", MX_C, c); return EAVSR_ERR_UNSUPPORTED; }
  EAVSR_REQUIRE(al16(a) && al16(b) && al16(out), "adapt_mix: tensors must be 16-byte aligned dense NHWC");
  EAVSR_REQUIRE(2 * n <= 65535 && ceil_div(h, MX_TH) <= 65535, "adapt_mix: batch/height too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == EAVSR_F32) return mix_launch<float>(a, b, w1, b1, w2, b2, out, n, h, w, negative_slope, st);
  if (dtype == EAVSR_BF16) return mix_launch<__nv_bfloat16>(a, b, w1, b1, w2, b2, out, n, h, w, negative_slope, st);
  set_error("adapt_mix: bad dtype %d", dtype);
  return EAVSR_ERR_INVALID;
}

extern "C" int eavsr_affine_offsets_forward(const void* transform, const int64_t transform_strides[4],
                                            const void* translation, const int64_t translation_strides[4],
                                            const void* mask_logits, const int64_t mask_strides[4],
                                            const void* transform_bias, const void* translation_bias,
                                            const void* mask_bias, float* offset, float* mask, int n,
                                            int deform_groups, int h, int w, int dtype, void* stream) {
  EAVSR_REQUIRE(transform && translation && offset && transform_strides && translation_strides,
                "affine_offsets: null pointer");
  EAVSR_REQUIRE(!mask || (mask_logits && mask_strides), "affine_offsets: mask output without logits");
  EAVSR_REQUIRE(n > 0 && deform_groups > 0 && h > 0 && w > 0, "affine_offsets: empty tensor");
  cudaStream_t st = (cudaStream_t)stream;
  Strides4 ts{transform_strides[0], transform_strides[1], transform_strides[2], transform_strides[3]};
  Strides4 rs{translation_strides[0], translation_strides[1], translation_strides[2], translation_strides[3]};
  Strides4 ms{0, 0, 0, 0};
  if (mask) ms = Strides4{mask_strides[0], mask_strides[1], mask_strides[2], mask_strides[3]};
  const long long total = (long long)n * deform_groups * h * w;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (dtype == EAVSR_F32)
    affine_offsets_kernel<float><<<blocks, 256, 0, st>>>((const float*)transform, ts, (const float*)translation, rs,
                                                         (const float*)mask_logits, ms, (const float*)transform_bias,
                                                         (const float*)translation_bias, (const float*)mask_bias,
                                                         offset, mask, n, deform_groups, h, w);
  else if (dtype == EAVSR_BF16)
    affine_offsets_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
        (const __nv_bfloat16*)transform, ts, (const __nv_bfloat16*)translation, rs,
        (const __nv_bfloat16*)mask_logits, ms, (const __nv_bfloat16*)transform_bias,
        (const __nv_bfloat16*)translation_bias, (const __nv_bfloat16*)mask_bias, offset, mask, n, deform_groups, h, w);
  else { set_error("affine_offsets: bad dtype %d", dtype); return EAVSR_ERR_INVALID; }
  return check_launch("affine_offsets");
}

extern "C" int eavsr_ca_residual_forward(const void* res, const void* skip, const void* w1, const void* b1,
                                         const void* w2, const void* b2, const void* res_bias, void* out,
                                         float* sums_workspace, int n, int c, int h, int w, int reduction, int dtype,
                                         void* stream) {
  EAVSR_REQUIRE(res && skip && w1 && b1 && w2 && b2 && out && sums_workspace, "ca_residual: null pointer");
  EAVSR_REQUIRE(n > 0 && h > 0 && w > 0, "ca_residual: empty tensor");
  if (c != 64 || reduction != 16) {
    set_error("ca_residual: only C=64, reduction=16 is fused (got C=%d, r=%d)", c, reduction);
    return EAVSR_ERR_UNSUPPORTED;
  }
  EAVSR_REQUIRE(al16(res) && al16(skip) && al16(out), "ca_residual: tensors must be 16-byte aligned dense NHWC");
  EAVSR_REQUIRE(n <= 65535, "ca_residual: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = h * w;
  cudaError_t e = cudaMemsetAsync(sums_workspace, 0, (size_t)n * c * sizeof(float), st);
  if (e != cudaSuccess) { set_error("ca_residual: memset: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
  dim3 g1(ceil_div(HW, CS_PIX), n);
  int rc;
  if (dtype == EAVSR_F32) {
    channel_sum_kernel<float, 64><<<g1, CS_THREADS, 0, st>>>((const float*)res, sums_workspace, HW);
    rc = check_launch("ca_residual(sum)");
    if (rc) return rc;
    const int chunks = HW * 64 / 4;
    dim3 g2(min(ceil_div(chunks, 256 * 4), 148 * 8), n);
    ca_scale_residual_kernel<float, 64, 4><<<g2, 256, 0, st>>>((const float*)res, (const float*)skip, sums_workspace,
                                                             (const float*)w1, (const float*)b1, (const float*)w2,
                                                             (const float*)b2, (const float*)res_bias, (float*)out, HW,
                                                             1.f / (float)HW, chunks);
  } else if (dtype == EAVSR_BF16) {
    using B = __nv_bfloat16;
    channel_sum_kernel<B, 64><<<g1, CS_THREADS, 0, st>>>((const B*)res, sums_workspace, HW);
    rc = check_launch("ca_residual(sum)");
    if (rc) return rc;
    const int chunks = HW * 64 / 8;
    dim3 g2(min(ceil_div(chunks, 256 * 4), 148 * 8), n);
    ca_scale_residual_kernel<B, 64, 4><<<g2, 256, 0, st>>>((const B*)res, (const B*)skip, sums_workspace, (const B*)w1,
                                                         (const B*)b1, (const B*)w2, (const B*)b2, (const B*)res_bias,
                                                         (B*)out, HW, 1.f / (float)HW, chunks);
  } else { set_error("ca_residual: bad dtype %d", dtype); return EAVSR_ERR_INVALID; }
  return check_launch("ca_residual(scale)");
}

extern "C" int eavsr_bias_act_forward(void* x, const void* bias, int c, long long pixels, float negative_slope,
                                      int dtype, void* stream) {
  EAVSR_REQUIRE(x && bias, "bias_act: null pointer");
  EAVSR_REQUIRE(c > 0 && pixels > 0, "bias_act: empty tensor");
  const int vec = dtype == EAVSR_F32 ? 4 : 8;
  if (c % vec != 0 || !al16(x) || !al16(bias)) {
    set_error("bias_act: needs C %% %d == 0 and 16-byte aligned dense NHWC data (C=%d)", vec, c);
    return EAVSR_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long nchunks = pixels * (c / vec);
  long long blocks = (nchunks + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype == EAVSR_F32)
    bias_act_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((float*)x, (const float*)bias, c, nchunks, negative_slope);
  else if (dtype == EAVSR_BF16)
    bias_act_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((__nv_bfloat16*)x, (const __nv_bfloat16*)bias, c,
                                                                    nchunks, negative_slope);
  else { set_error("bias_act: bad dtype %d", dtype); return EAVSR_ERR_INVALID; }
  return check_launch("bias_act");
}

// ---------------------------------------------------------------------------------------------
// nhwc_cat: torch.cat(dim=1) of dense NHWC tensors into (a channel slice of) an NHWC buffer
// ---------------------------------------------------------------------------------------------
// The concatenations in front of the fusion / backbone / reconstruction convolutions (models/eavsrp_model.py:
// 271-324, 350-364; 128..320 channels at full resolution) ran at ~1 TB/s in ATen's generic cat kernel (it does
// not know the tensors are channels_last) and were 6.5 % of the device time.  Here every thread moves 16-byte
// chunks: consecutive threads read consecutive chunks of one source pixel and write consecutive chunks of the
// output pixel, both fully coalesced.
constexpr int CAT_MAX = 8;
struct CatParams {
  const uint4* src[CAT_MAX];
  int cpp[CAT_MAX];          // 16-byte chunks per pixel of each source
  int first[CAT_MAX + 1];    // prefix sums of cpp
  int nsrc;
  int out_cpp;               // chunks per pixel of the OUTPUT buffer (its pixel stride)
  int out_off;               // first chunk written inside an output pixel
};

__global__ void __launch_bounds__(256)
nhwc_cat_kernel(const __grid_constant__ CatParams P, uint4* __restrict__ out, long long pixels) {
  const int total_cpp = P.first[P.nsrc];
  const long long n = pixels * total_cpp;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float rcp = 1.f / (float)total_cpp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long long px;
    int ch;
    if (n < (1ll << 22)) {                       // float reciprocal + one correction: exact below 2^22
      int q = __float2int_rz((float)(int)i * rcp), r = (int)i - q * total_cpp;
      if (r < 0) { r += total_cpp; --q; }
      if (r >= total_cpp) { r -= total_cpp; ++q; }
      px = q; ch = r;
    } else {
      px = i / total_cpp; ch = (int)(i - px * total_cpp);
    }
    int s = 0;
#pragma unroll
    for (int k = 1; k < CAT_MAX; ++k) s += (k < P.nsrc && ch >= P.first[k]) ? 1 : 0;
    const uint4 v = __ldg(P.src[s] + px * P.cpp[s] + (ch - P.first[s]));
    out[px * P.out_cpp + P.out_off + ch] = v;
  }
}

// Scale + residual only, with channel sums that were produced elsewhere (the epilogue of
// eavsr_conv3x3_forward): out = (res + res_bias) * sigmoid(MLP(sums / HW + res_bias)) + skip.
extern "C" int eavsr_ca_scale_forward(const void* res, const void* skip, const float* sums, const void* w1,
                                      const void* b1, const void* w2, const void* b2, const void* res_bias, void* out,
                                      int n, int c, int h, int w, int reduction, int dtype, void* stream) {
  EAVSR_REQUIRE(res && skip && sums && w1 && b1 && w2 && b2 && out, "ca_scale: null pointer");
  EAVSR_REQUIRE(n > 0 && h > 0 && w > 0 && n <= 65535, "ca_scale: bad shape");
  if (c != 64 || reduction != 16) {
    set_error("ca_scale: only C=64, reduction=16 is fused (got C=%d, r=%d)", c, reduction);
    return EAVSR_ERR_UNSUPPORTED;
  }
  EAVSR_REQUIRE(al16(res) && al16(skip) && al16(out), "ca_scale: tensors must be 16-byte aligned dense NHWC");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = h * w;
  if (dtype == EAVSR_F32) {
    const int chunks = HW * 64 / 4;
    dim3 g2(min(ceil_div(chunks, 256 * 4), 148 * 8), n);
    ca_scale_residual_kernel<float, 64, 4><<<g2, 256, 0, st>>>((const float*)res, (const float*)skip, sums,
                                                             (const float*)w1, (const float*)b1, (const float*)w2,
                                                             (const float*)b2, (const float*)res_bias, (float*)out, HW,
                                                             1.f / (float)HW, chunks);
  } else if (dtype == EAVSR_BF16) {
    using B = __nv_bfloat16;
    const int chunks = HW * 64 / 8;
    dim3 g2(min(ceil_div(chunks, 256 * 4), 148 * 8), n);
    ca_scale_residual_kernel<B, 64, 4><<<g2, 256, 0, st>>>((const B*)res, (const B*)skip, sums, (const B*)w1,
                                                         (const B*)b1, (const B*)w2, (const B*)b2, (const B*)res_bias,
                                                         (B*)out, HW, 1.f / (float)HW, chunks);
  } else { set_error("ca_scale: bad dtype %d", dtype); return EAVSR_ERR_INVALID; }
  return check_launch("ca_scale");
}

extern "C" int eavsr_bias_act_shuffle_forward(const void* x, const void* bias, void* out, int n, int c_out, int h,
                                              int w, float negative_slope, int dtype, void* stream) {
  EAVSR_REQUIRE(x && out, "bias_act_shuffle: null pointer");
  EAVSR_REQUIRE(n > 0 && c_out > 0 && h > 0 && w > 0, "bias_act_shuffle: empty tensor");
  const int vec = dtype == EAVSR_F32 ? 4 : 8;
  if (c_out % vec != 0 || !al16(x) || !al16(out) || (bias && !al16(bias))) {
    set_error("bias_act_shuffle: needs C_out %% %d == 0 and 16-byte aligned dense NHWC data (C_out=%d)", vec, c_out);
    return EAVSR_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long items = (long long)n * h * w * (c_out / vec);
  long long blocks = (items + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype == EAVSR_F32)
    bias_act_shuffle_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)x, (const float*)bias, (float*)out,
                                                                    c_out, h, w, items, negative_slope);
  else if (dtype == EAVSR_BF16)
    bias_act_shuffle_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)bias, (__nv_bfloat16*)out, c_out, h, w, items, negative_slope);
  else { set_error("bias_act_shuffle: bad dtype %d", dtype); return EAVSR_ERR_INVALID; }
  return check_launch("bias_act_shuffle");
}

extern "C" int eavsr_nhwc_cat_forward(const void* const* srcs, const int* src_channels, int nsrc, void* out,
                                      int out_channels, int out_channel_offset, long long pixels, int dtype,
                                      void* stream) {
  EAVSR_REQUIRE(srcs && src_channels && out, "nhwc_cat: null pointer");
  EAVSR_REQUIRE(nsrc >= 1 && nsrc <= CAT_MAX, "nhwc_cat: 1..%d sources (got %d)", CAT_MAX, nsrc);
  EAVSR_REQUIRE(pixels > 0, "nhwc_cat: empty tensor");
  if (dtype != EAVSR_F32 && dtype != EAVSR_BF16) { set_error("nhwc_cat: bad dtype %d", dtype); return EAVSR_ERR_INVALID; }
  const int vec = dtype == EAVSR_F32 ? 4 : 8;
  CatParams P;
  P.nsrc = nsrc;
  P.first[0] = 0;
  for (int i = 0; i < nsrc; ++i) {
    EAVSR_REQUIRE(srcs[i], "nhwc_cat: null source %d", i);
    if (src_channels[i] <= 0 || src_channels[i] % vec != 0 || !al16(srcs[i])) {
      set_error("nhwc_cat: source %d needs C %% %d == 0 and 16-byte alignment (C=%d)", i, vec, src_channels[i]);
      return EAVSR_ERR_UNSUPPORTED;
    }
    P.src[i] = (const uint4*)srcs[i];
    P.cpp[i] = src_channels[i] / vec;
    P.first[i + 1] = P.first[i] + P.cpp[i];
  }
  for (int i = nsrc; i < CAT_MAX; ++i) { P.src[i] = nullptr; P.cpp[i] = 0; P.first[i + 1] = P.first[nsrc]; }
  if (out_channels % vec != 0 || out_channel_offset % vec != 0 || out_channel_offset < 0 || !al16(out) ||
      out_channel_offset / vec + P.first[nsrc] > out_channels / vec) {
    set_error("nhwc_cat: output slice [%d, +%d) of %d channels must be 16-byte aligned and inside the buffer",
              out_channel_offset, P.first[nsrc] * vec, out_channels);
    return EAVSR_ERR_INVALID;
  }
  P.out_cpp = out_channels / vec;
  P.out_off = out_channel_offset / vec;
  const long long n = pixels * P.first[nsrc];
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  nhwc_cat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(P, (uint4*)out, pixels);
  return check_launch("nhwc_cat");
}

// Per-(n, channel) sums over H*W of a dense NHWC tensor with 64 channels, fp32 (sums is zero-filled here): the
// bias gradient of the 64-output convolutions and the global average pooling of CALayer in the training step, where
// ATen's generic reduction of a channels_last tensor took 30 us per 4 MB tensor.
extern "C" int eavsr_channel_sum_forward(const void* x, float* sums, int n, int c, long long hw, int dtype, void* stream) {
  EAVSR_REQUIRE(x && sums, "channel_sum: null pointer");
  EAVSR_REQUIRE(n > 0 && hw > 0 && hw < (1ll << 31) && n <= 65535, "channel_sum: bad shape");
  if (c != 64) { set_error("channel_sum: only C=64 (got %d)", c); return EAVSR_ERR_UNSUPPORTED; }
  EAVSR_REQUIRE(al16(x), "channel_sum: x must be 16-byte aligned dense NHWC");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)n * c * sizeof(float), st);
  if (e != cudaSuccess) { set_error("channel_sum: memset: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
  dim3 g1(ceil_div(hw, CS_PIX), n);
  if (dtype == EAVSR_F32) channel_sum_kernel<float, 64><<<g1, CS_THREADS, 0, st>>>((const float*)x, sums, (int)hw);
  else if (dtype == EAVSR_BF16) channel_sum_kernel<__nv_bfloat16, 64><<<g1, CS_THREADS, 0, st>>>((const __nv_bfloat16*)x, sums, (int)hw);
  else { set_error("channel_sum: bad dtype %d", dtype); return EAVSR_ERR_INVALID; }
  return check_launch("channel_sum");
}


// sums(n, 64) = sum over the pixels of a * b (both (n, 64, h, w) dense NHWC): the gradient of the channel-attention
// scale in `res * scale + skip` (d scale = sum_hw(grad * res)), one pass instead of a multiply + ATen reduction.
extern "C" int eavsr_channel_dot_forward(const void* a, const void* b, float* sums, int n, int c, long long hw, int dtype,
                                         void* stream) {
  EAVSR_REQUIRE(a && b && sums, "channel_dot: null pointer");
  EAVSR_REQUIRE(n > 0 && hw > 0 && hw < (1ll << 31) && n <= 65535, "channel_dot: bad shape");
  if (c != 64) { set_error("channel_dot: only C=64 (got %d)", c); return EAVSR_ERR_UNSUPPORTED; }
  EAVSR_REQUIRE(al16(a) && al16(b), "channel_dot: tensors must be 16-byte aligned dense NHWC");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)n * c * sizeof(float), st);
  if (e != cudaSuccess) { set_error("channel_dot: memset: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
  dim3 g1(ceil_div(hw, CS_PIX), n);
  if (dtype == EAVSR_F32)
    channel_sum_kernel<float, 64, true><<<g1, CS_THREADS, 0, st>>>((const float*)a, sums, (int)hw, (const float*)b);
  else if (dtype == EAVSR_BF16)
    channel_sum_kernel<__nv_bfloat16, 64, true><<<g1, CS_THREADS, 0, st>>>((const __nv_bfloat16*)a, sums, (int)hw,
                                                                          (const __nv_bfloat16*)b);
  else { set_error("channel_dot: bad dtype %d", dtype); return EAVSR_ERR_INVALID; }
  return check_launch("channel_dot");
}
