// Differentiable grouped 3x3 convolutions of AdaptBlockOffset / AdaptBlock2_3x3 (SURVEY.md section 8 row a3,
// models/networks.py:289-290, 327-328): `concat` = depthwise 3x3 over the 128 concatenated channels and `concat2` =
// grouped 3x3 with two input channels per output channel (128 -> 64, groups = 64), stride 1, padding 1.
//
// Inference fuses both (and their LeakyReLUs) in adapt_mix_kernel.  The TRAINING step ran them through cuDNN, whose
// grouped kernels are the single largest item of the step: `convolution_backward` of the 128 -> 64 grouped conv takes
// 647 us per call on 8 x 64 x 64 crops (216 calls = 14.6 % of the step), plus its layout transforms
// (`tensorTransformGeneric`, 13.7 %) and the grouped-direct forward (4.6 %) -- for 8 MB tensors that a bandwidth-bound
// kernel moves in a few microseconds.  Forward, d(input) and d(weight) + d(bias) here are plain NHWC SIMT kernels
// (16-byte chunks, fp32 accumulation); inputs per group CPG in {1, 2}, one output channel per group.
#include "common.cuh"

namespace eavsr {
namespace {

constexpr int GC_THREADS = 256;
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T> struct Chunk8;     // 8 consecutive channels <-> fp32[8]
template <> struct Chunk8<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* f) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    f[0] = bf16lo_to_f32(u.x); f[1] = bf16hi_to_f32(u.x); f[2] = bf16lo_to_f32(u.y); f[3] = bf16hi_to_f32(u.y);
    f[4] = bf16lo_to_f32(u.z); f[5] = bf16hi_to_f32(u.z); f[6] = bf16lo_to_f32(u.w); f[7] = bf16hi_to_f32(u.w);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* f) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct Chunk8<float> {
  static __device__ __forceinline__ void ld(const float* p, float* f) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* f) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// forward: one thread = one pixel x 8 OUTPUT channels (= 8 * CPG input channels)
template <typename T, int CPG>
__global__ void __launch_bounds__(GC_THREADS)
gconv_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ b, T* __restrict__ out, int N,
                 int H, int W, int Cout) {
  const int cpo = Cout / 8;                                   // output chunks per pixel
  const long long total = (long long)N * H * W * cpo;
  const int Cin = Cout * CPG;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cpo);
    const long long px = i / cpo;
    const int xq = (int)(px % W), y = (int)((px / W) % H);
    const long long n = px / ((long long)W * H);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = b ? to_f32<T>(b[ch * 8 + e]) : 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y + t / 3 - 1, xx = xq + t % 3 - 1;
      if ((unsigned)yy >= (unsigned)H || (unsigned)xx >= (unsigned)W) continue;
      const T* src = x + ((n * H + yy) * (long long)W + xx) * Cin + ch * 8 * CPG;
      float v[8 * CPG];
      Chunk8<T>::ld(src, v);
      if (CPG == 2) Chunk8<T>::ld(src + 8, v + 8);
#pragma unroll
      for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int j = 0; j < CPG; ++j)                         // weight (cout, CPG, 3, 3)
          acc[e] += v[e * CPG + j] * to_f32<T>(w[((ch * 8 + e) * CPG + j) * 9 + t]);
    }
    Chunk8<T>::st(out + px * Cout + ch * 8, acc);
  }
}

// d(input): one thread = one pixel x 8 INPUT channels (= 8 / CPG output channels of gout); the transposed stencil
template <typename T, int CPG>
__global__ void __launch_bounds__(GC_THREADS)
gconv_bwd_data_kernel(const T* __restrict__ g, const T* __restrict__ w, T* __restrict__ gx, int N, int H, int W, int Cout) {
  const int Cin = Cout * CPG, cpi = Cin / 8;
  const long long total = (long long)N * H * W * cpi;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cpi);                            // input chunk: channels 8*ch .. 8*ch+7
    const long long px = i / cpi;
    const int xq = (int)(px % W), y = (int)((px / W) % H);
    const long long n = px / ((long long)W * H);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y - (t / 3 - 1), xx = xq - (t % 3 - 1);   // the output pixel this input fed through tap t
      if ((unsigned)yy >= (unsigned)H || (unsigned)xx >= (unsigned)W) continue;
      const T* src = g + ((n * H + yy) * (long long)W + xx) * Cout;
      if (CPG == 1) {
        float v[8];
        Chunk8<T>::ld(src + ch * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v[e] * to_f32<T>(w[(ch * 8 + e) * 9 + t]);
      } else {                                                // inputs 8ch..8ch+7 <-> outputs 4ch..4ch+3
        float v[8];
        Chunk8<T>::ld(src + (ch >> 1) * 8, v);                // the 8-channel gout chunk holding outputs 4ch..4ch+3
        const bool hi = (ch & 1) != 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int co = ch * 4 + (e >> 1);
          acc[e] += (hi ? v[4 + (e >> 1)] : v[e >> 1]) * to_f32<T>(w[(co * 2 + (e & 1)) * 9 + t]);
        }
      }
    }
    Chunk8<T>::st(gx + px * Cin + ch * 8, acc);
  }
}

// d(weight), d(bias): thread (pixel lane, 8-input-channel chunk) accumulates its 9 x 8 weight gradients (and the
// matching output-channel sums) over a strip of pixels, the CTA reduces over its pixel lanes in shared memory and
// issues one atomicAdd per weight.  gw: (cout, CPG, 3, 3) fp32, gb: (cout) fp32, both ZERO on entry.
template <typename T, int CPG>
__global__ void __launch_bounds__(GC_THREADS)
gconv_bwd_weight_kernel(const T* __restrict__ g, const T* __restrict__ x, float* __restrict__ gw, float* __restrict__ gb,
                        int N, int H, int W, int Cout) {
  const int Cin = Cout * CPG, cpi = Cin / 8;                  // 16 chunks for 128 input channels
  const int lanes = GC_THREADS / cpi;                         // pixel lanes per CTA
  const int ch = threadIdx.x % cpi, pl = threadIdx.x / cpi;
  const long long pixels = (long long)N * H * W;
  float acc[9][8], accb[8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[t][e] = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) accb[e] = 0.f;
  if (pl < lanes) {
    for (long long px = (long long)blockIdx.x * lanes + pl; px < pixels; px += (long long)gridDim.x * lanes) {
      const int xq = (int)(px % W), y = (int)((px / W) % H);
      const long long n = px / ((long long)W * H);
      // gradient of the outputs fed by input channels 8ch..8ch+7 at THIS pixel, expanded per input channel
      float gv[8];
      if (CPG == 1) {
        Chunk8<T>::ld(g + px * Cout + ch * 8, gv);
      } else {
        float v[8];
        Chunk8<T>::ld(g + px * Cout + (ch >> 1) * 8, v);
        const bool hi = (ch & 1) != 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) gv[e] = hi ? v[4 + (e >> 1)] : v[e >> 1];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) accb[e] += gv[e];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = xq + t % 3 - 1;
        if ((unsigned)yy >= (unsigned)H || (unsigned)xx >= (unsigned)W) continue;
        float v[8];
        Chunk8<T>::ld(x + ((n * H + yy) * (long long)W + xx) * Cin + ch * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[t][e] += gv[e] * v[e];
      }
    }
  }
  // Reduce over the pixel lanes.  Threads with the same channel chunk are `cpi` apart: when cpi == 16 (128 input
  // channels) lanes l and l + 16 of a warp share a chunk, so one shuffle halves the work; the warps then meet in
  // shared memory one tap (8 values per thread) at a time -- 10 rounds instead of one per value.
  __shared__ float red[GC_THREADS / 32][32][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool fold = (cpi == 16);                               // lane and lane ^ 16 hold the same chunk
#pragma unroll
  for (int t = 0; t < 10; ++t) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = t < 9 ? acc[t][e] : accb[e];
      if (fold) v[e] += __shfl_xor_sync(0xffffffffu, v[e], 16);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[warp][lane][e] = v[e];
    __syncthreads();
    // one thread per (chunk, element): sum the contributions of every warp (and, without the fold, of every lane
    // of a warp that holds this chunk)
    for (int o = threadIdx.x; o < cpi * 8; o += GC_THREADS) {
      const int c8 = o >> 3, e = o & 7;
      float sum = 0.f;
      for (int wp = 0; wp < GC_THREADS / 32; ++wp) {
        if (fold) {
          sum += red[wp][c8][e];
        } else {
          for (int l = 0; l < 32; ++l)
            if ((wp * 32 + l) % cpi == c8 && (wp * 32 + l) / cpi < lanes) sum += red[wp][l][e];
        }
      }
      const int c = c8 * 8 + e;                               // input channel
      if (t < 9) atomicAdd(gw + c * 9 + t, sum);              // (cout, CPG, 3, 3) flattened == input channel * 9 + tap
      else if (CPG == 1 || (c & 1) == 0) atomicAdd(gb + (CPG == 1 ? c : c >> 1), sum);   // CPG = 2: one sum per group
    }
    __syncthreads();
  }
}

template <typename T, int CPG>
int gconv_run(int mode, const void* a, const void* b, const void* c, void* o1, float* o2, float* o3, int n, int h, int w,
              int cout, cudaStream_t st) {
  const long long chunks = (long long)n * h * w * (mode == 1 ? cout * CPG / 8 : cout / 8);
  long long blocks = (chunks + GC_THREADS - 1) / GC_THREADS;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (mode == 0) {
    gconv_fwd_kernel<T, CPG><<<(unsigned)blocks, GC_THREADS, 0, st>>>((const T*)a, (const T*)b, (const T*)c, (T*)o1, n, h, w, cout);
    return check_launch("grouped_conv3x3_forward");
  }
  if (mode == 1) {
    gconv_bwd_data_kernel<T, CPG><<<(unsigned)blocks, GC_THREADS, 0, st>>>((const T*)a, (const T*)b, (T*)o1, n, h, w, cout);
    return check_launch("grouped_conv3x3_backward(data)");
  }
  const int lanes = GC_THREADS / (cout * CPG / 8);
  long long wb = ((long long)n * h * w + lanes * 16 - 1) / (lanes * 16);   // >= 16 pixels per thread
  if (wb > 148 * 2) wb = 148 * 2;
  if (wb < 1) wb = 1;
  gconv_bwd_weight_kernel<T, CPG><<<(unsigned)wb, GC_THREADS, 0, st>>>((const T*)a, (const T*)b, o2, o3, n, h, w, cout);
  return check_launch("grouped_conv3x3_backward(weight)");
}

int gconv_dispatch(int mode, const void* a, const void* b, const void* c, void* o1, float* o2, float* o3, int n, int cin,
                   int cout, int h, int w, int dtype, cudaStream_t st, const char* who) {
  EAVSR_REQUIRE(n > 0 && h > 0 && w > 0, "%s: empty tensor", who);
  const int cpg = cout > 0 ? cin / cout : 0;
  if (cout <= 0 || cin != cout * cpg || (cpg != 1 && cpg != 2) || cout % 8 != 0 || cin % 8 != 0 || GC_THREADS % (cin / 8) != 0 ||
      cin / 8 > GC_THREADS) {
    set_error("%s: only groups = cout with 1 or 2 inputs per group and channel counts that are multiples of 8 "
              "(got %d -> %d)", who, cin, cout);
    return EAVSR_ERR_UNSUPPORTED;
  }
  if (dtype == EAVSR_BF16)
    return cpg == 1 ? gconv_run<__nv_bfloat16, 1>(mode, a, b, c, o1, o2, o3, n, h, w, cout, st)
                    : gconv_run<__nv_bfloat16, 2>(mode, a, b, c, o1, o2, o3, n, h, w, cout, st);
  if (dtype == EAVSR_F32)
    return cpg == 1 ? gconv_run<float, 1>(mode, a, b, c, o1, o2, o3, n, h, w, cout, st)
                    : gconv_run<float, 2>(mode, a, b, c, o1, o2, o3, n, h, w, cout, st);
  set_error("%s: bad dtype %d", who, dtype);
  return EAVSR_ERR_INVALID;
}

}  // namespace
}  // namespace eavsr

using namespace eavsr;

extern "C" int eavsr_grouped_conv3x3_forward(const void* x, const void* weight, const void* bias, void* out, int n, int cin,
                                             int cout, int h, int w, int dtype, void* stream) {
  EAVSR_REQUIRE(x && weight && out, "grouped_conv3x3_forward: null pointer");
  EAVSR_REQUIRE(al16(x) && al16(out), "grouped_conv3x3_forward: tensors must be 16-byte aligned dense NHWC");
  return gconv_dispatch(0, x, weight, bias, out, nullptr, nullptr, n, cin, cout, h, w, dtype, (cudaStream_t)stream,
                        "grouped_conv3x3_forward");
}

extern "C" int eavsr_grouped_conv3x3_backward(const void* gout, const void* x, const void* weight, void* gx, float* gweight,
                                              float* gbias, int n, int cin, int cout, int h, int w, int dtype,
                                              void* stream) {
  EAVSR_REQUIRE(gout && weight && (!gweight || x), "grouped_conv3x3_backward: null pointer");
  EAVSR_REQUIRE(al16(gout) && (!x || al16(x)) && (!gx || al16(gx)),
                "grouped_conv3x3_backward: tensors must be 16-byte aligned dense NHWC");
  EAVSR_REQUIRE((gweight != nullptr) == (gbias != nullptr), "grouped_conv3x3_backward: gweight and gbias go together");
  int rc = EAVSR_OK;
  if (gx)
    rc = gconv_dispatch(1, gout, weight, nullptr, gx, nullptr, nullptr, n, cin, cout, h, w, dtype, (cudaStream_t)stream,
                        "grouped_conv3x3_backward");
  if (rc == EAVSR_OK && gweight)
    rc = gconv_dispatch(2, gout, x, nullptr, nullptr, gweight, gbias, n, cin, cout, h, w, dtype, (cudaStream_t)stream,
                        "grouped_conv3x3_backward");
  return rc;
}
