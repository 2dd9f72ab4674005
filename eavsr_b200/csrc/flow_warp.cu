// flow_warp forward / backward for sm_100a.
//
// Replaces the reference's flow_warp (models/networks.py:699-739, models/eavsrp_model.py:587-626):
// host-side meshgrid + normalise + F.grid_sample(bilinear, align_corners=True).  Because
// align_corners=True maps the normalised grid back to exactly (x + flow_x, y + flow_y) the kernel
// samples at that pixel coordinate directly; no grid tensor ever exists.
//
// HBM-bound gather.  Fast path: NHWC features, one thread per (pixel, 16-byte channel chunk),
// 8x16-pixel CTA tiles so that the 4 bilinear corners of neighbouring outputs hit L1, all four
// 16-byte corner loads of a thread in flight together.  Algorithmic bytes per pixel:
// 2*C*sizeof(T) + 8 (read x once, read flow, write out) -- see DESIGN.md.
#include <type_traits>

#include "common.cuh"

namespace eavsr {

namespace {

constexpr int TILE_H = 8;
constexpr int TILE_W = 16;
constexpr int TILE_PIX = TILE_H * TILE_W;
constexpr int WARP_THREADS = 256;

struct Corner {
  int off[4];    // pixel index (y*W+x) of the 4 corners, clamped in-range
  float wgt[4];  // bilinear weight, 0 for out-of-image corners
  float gxm, gym;  // d(coord)/d(flow): 0 when the border clamp is active
  float ly, lx;
  bool ok[4];
};

template <int PAD>
__device__ __forceinline__ Corner make_corner(float sy, float sx, int H, int W) {
  Corner c;
  c.gxm = 1.f;
  c.gym = 1.f;
  if (PAD == EAVSR_PAD_BORDER) {
    // grid_sampler clip_coordinates_set_grad: gradient 0 when in<=0 or in>=size-1.
    if (sx <= 0.f) { sx = 0.f; c.gxm = 0.f; } else if (sx >= (float)(W - 1)) { sx = (float)(W - 1); c.gxm = 0.f; }
    if (sy <= 0.f) { sy = 0.f; c.gym = 0.f; } else if (sy >= (float)(H - 1)) { sy = (float)(H - 1); c.gym = 0.f; }
  }
  // keep the int conversion defined for wild flows
  sx = fminf(fmaxf(sx, -2.f), (float)W + 1.f);
  sy = fminf(fmaxf(sy, -2.f), (float)H + 1.f);
  float fy = floorf(sy), fx = floorf(sx);
  int y0 = (int)fy, x0 = (int)fx;
  float ly = sy - fy, lx = sx - fx;
  c.ly = ly;
  c.lx = lx;
  int y1 = y0 + 1, x1 = x0 + 1;
  bool vy0 = (y0 >= 0) && (y0 < H), vy1 = (y1 >= 0) && (y1 < H);
  bool vx0 = (x0 >= 0) && (x0 < W), vx1 = (x1 >= 0) && (x1 < W);
  int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
  int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
  c.ok[0] = vy0 && vx0; c.ok[1] = vy0 && vx1; c.ok[2] = vy1 && vx0; c.ok[3] = vy1 && vx1;
  c.off[0] = cy0 * W + cx0; c.off[1] = cy0 * W + cx1; c.off[2] = cy1 * W + cx0; c.off[3] = cy1 * W + cx1;
  c.wgt[0] = c.ok[0] ? (1.f - ly) * (1.f - lx) : 0.f;
  c.wgt[1] = c.ok[1] ? (1.f - ly) * lx : 0.f;
  c.wgt[2] = c.ok[2] ? ly * (1.f - lx) : 0.f;
  c.wgt[3] = c.ok[3] ? ly * lx : 0.f;
  return c;
}

__device__ __forceinline__ void load_flow(const float* __restrict__ flow, int layout, int n, int y, int x, int H,
                                          int W, float& fx, float& fy) {
  if (layout == EAVSR_FLOW_N2HW) {
    size_t b = ((size_t)n * 2) * H * W + (size_t)y * W + x;
    fx = __ldg(flow + b);
    fy = __ldg(flow + b + (size_t)H * W);
  } else {
    float2 f = __ldg(reinterpret_cast<const float2*>(flow) + ((size_t)n * H + y) * W + x);
    fx = f.x;
    fy = f.y;
  }
  // the reference normalises with 2 / max(size - 1, 1) (models/networks.py:728-729) and grid_sample maps back
  // with (size - 1) / 2: along a size-1 dimension the sample coordinate is 0 whatever the flow says
  if (W == 1) fx = 0.f;
  if (H == 1) fy = 0.f;
}

// ------------------------------------------------------------------------------------------
// Lean fast path: NHWC, power-of-two chunks per pixel (CPP = C*sizeof(T)/16 in {1,...,32}).
//
// The first version of this kernel was instruction-issue bound, not HBM bound (ncu: 629 warp
// instructions per 4 pixels, 57 % issue-active, 11 % DRAM): with (pixel, chunk) threads the 8 lanes of
// a pixel all redo the same coordinate / index arithmetic.  Here a warp owns a 4x8-pixel patch:
// phase 1, lane <-> pixel computes the clamped corner base, the +x / +row steps and the four
// weights ONCE; phase 2 walks the patch 32/CPP pixels at a time, lane <-> (pixel, 16-byte chunk),
// fetching those 7 values with warp shuffles.  Loads are unconditional (out-of-image corners read
// a clamped in-image address with weight 0), offsets are 32-bit.  CTA = 4 warps = an 8x16 tile.
// Two documented divergences from F.grid_sample, both outside anything EAVSR produces: (i) an Inf / NaN
// feature value next to the image border can reach an out-of-image corner's load (0 * Inf = NaN) where
// grid_sample never reads it -- the general kernels below select per corner instead; (ii) a NaN flow is
// clamped to the out-of-image sentinel and yields 0 (ATen's float->int cast of NaN is undefined).
// ------------------------------------------------------------------------------------------
constexpr int LEAN_THREADS = 128;

// Pyramid flows: MultiAdSTN.forward (models/networks.py:600-615) warps with flows that are
// sums of bilinearly resized (align_corners=True), rescaled coarser / finer flow fields --
//   level 3:  interp(offset, 1/4) / 4;   level 2:  interp(offset, 1/2) / 2 + interp(p1, 2) * 2;
//   level 1:  offset + interp(p2 + p1_up, 2) * 2;    final:  p3 + p2_up + offset
// -- each materialised by F.interpolate + an elementwise launch or two in the reference.  Here the lane <-> pixel
// phase of the warp evaluates  flow(y, x) = sum_i scale_i * resize(flow_i)(y, x)  itself, with ATen's
// upsample_bilinear2d arithmetic (source index = dst * (in - 1) / (out - 1), lambda = fraction); a term can also be
// written out (`scaled_out`: level 2 keeps p1_up for level 1), and so can the sum (`flow_out`).
constexpr int PYR_MAX = 4;
struct PyrTerm {
  const float* flow;   // (n, 2, h, w) fp32 contiguous
  float* scaled_out;   // optional (n, 2, H, W): scale * resize(flow)
  int h, w;
  float scale, ry, rx; // ry = (h - 1) / (H - 1) (0 when H == 1)
};
struct PyrParams {
  int nterms;
  float* flow_out;     // optional (n, 2, H, W): the sum
  PyrTerm t[PYR_MAX];
};

__device__ __forceinline__ void pyr_eval(const PyrParams& P, int n, int y, int x, int H, int W, float& fx, float& fy) {
  fx = 0.f;
  fy = 0.f;
  const size_t opix = (size_t)y * W + x, oplane = (size_t)H * W;
#pragma unroll
  for (int i = 0; i < PYR_MAX; ++i) {
    if (i < P.nterms) {
      const PyrTerm& t = P.t[i];
      float vx, vy;
      const float* f0 = t.flow + (size_t)n * 2 * t.h * t.w;
      const size_t pl = (size_t)t.h * t.w;
      if (t.h == H && t.w == W) {
        vx = __ldg(f0 + opix);
        vy = __ldg(f0 + pl + opix);
      } else {
        const float sy = t.ry * (float)y, sx = t.rx * (float)x;
        const int y0 = (int)sy, x0 = (int)sx;
        const int yp = (y0 < t.h - 1) ? t.w : 0, xp = (x0 < t.w - 1) ? 1 : 0;
        const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
        const float* p = f0 + (size_t)y0 * t.w + x0;
        vx = hy * (hx * __ldg(p) + lx * __ldg(p + xp)) + ly * (hx * __ldg(p + yp) + lx * __ldg(p + yp + xp));
        p += pl;
        vy = hy * (hx * __ldg(p) + lx * __ldg(p + xp)) + ly * (hx * __ldg(p + yp) + lx * __ldg(p + yp + xp));
      }
      vx *= t.scale;
      vy *= t.scale;
      if (t.scaled_out) {
        t.scaled_out[(size_t)n * 2 * oplane + opix] = vx;
        t.scaled_out[(size_t)n * 2 * oplane + oplane + opix] = vy;
      }
      fx += vx;
      fy += vy;
    }
  }
  if (P.flow_out) {
    P.flow_out[(size_t)n * 2 * oplane + opix] = fx;
    P.flow_out[(size_t)n * 2 * oplane + oplane + opix] = fy;
  }
  if (W == 1) fx = 0.f;
  if (H == 1) fy = 0.f;
}

// DUAL: two feature maps warped with the same flow in one launch (nbr and feat_prop
// in MultiAdSTN.forward, models/networks.py:621-623) -- the flow read and phase 1 are shared.
struct NoPyr {};
template <typename T, int CPP, int PAD, bool DUAL = false, bool PYR = false>
__global__ void __launch_bounds__(LEAN_THREADS)
flow_warp_fwd_lean(const T* __restrict__ x, const float* __restrict__ flow, T* __restrict__ out, int H, int W,
                   int layout, int tiles_x, int tiles_y, long long xs_n, long long os_n,
                   const T* __restrict__ x2 = nullptr, T* __restrict__ out2 = nullptr, long long xs2_n = 0,
                   long long os2_n = 0, const __grid_constant__ typename std::conditional<PYR, PyrParams, NoPyr>::type pyr = {}) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int C = CPP * VEC;
  constexpr int PPI = 32 / CPP;          // pixels per phase-2 iteration
  int tile = blockIdx.x;
  const int n = tile / (tiles_x * tiles_y);
  tile -= n * tiles_x * tiles_y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // warp patch origin inside the 8x16 CTA tile (2x2 patches of 4 rows x 8 cols)
  const int py0 = (tile / tiles_x) * TILE_H + (warp >> 1) * 4;
  const int px0 = (tile % tiles_x) * TILE_W + (warp & 1) * 8;
  const T* xn = x + (size_t)n * xs_n;
  T* on = out + (size_t)n * os_n;

  // phase 1: lane <-> pixel (row = lane / 8, col = lane % 8)
  const int y = py0 + (lane >> 3), xq = px0 + (lane & 7);
  const bool live = (y < H) && (xq < W);
  uint32_t base = 0, stepx = 0, stepy = 0;
  float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
  if (live) {
    float fx, fy;
    if constexpr (PYR) pyr_eval(pyr, n, y, xq, H, W, fx, fy);
    else load_flow(flow, layout, n, y, xq, H, W, fx, fy);
    float sx = (float)xq + fx, sy = (float)y + fy;
    if (PAD == EAVSR_PAD_BORDER) {
      sx = fminf(fmaxf(sx, 0.f), (float)(W - 1));
      sy = fminf(fmaxf(sy, 0.f), (float)(H - 1));
    }
    sx = fminf(fmaxf(sx, -2.f), (float)W + 1.f);
    sy = fminf(fmaxf(sy, -2.f), (float)H + 1.f);
    const float fyf = floorf(sy), fxf = floorf(sx);
    const int y0 = (int)fyf, x0 = (int)fxf;
    const float ly = sy - fyf, lx = sx - fxf;
    const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
    const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
    const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
    const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1);
    base = (uint32_t)(cy0 * W + cx0) * C;
    stepx = (uint32_t)(cx1 - cx0) * C;
    stepy = (uint32_t)((cy1 - cy0) * W) * C;
    w00 = (vy0 && vx0) ? (1.f - ly) * (1.f - lx) : 0.f;
    w01 = (vy0 && vx1) ? (1.f - ly) * lx : 0.f;
    w10 = (vy1 && vx0) ? ly * (1.f - lx) : 0.f;
    w11 = (vy1 && vx1) ? ly * lx : 0.f;
  }
  const uint32_t live_mask = __ballot_sync(0xffffffffu, live);

  // phase 2: lane <-> (pixel sub*PPI + lane / CPP, chunk lane % CPP).  UN sub-steps are batched so
  // that 4*UN 16-byte loads per thread are in flight before the first blend (the un-batched
  // version serialised CPP round trips to L2/HBM per warp).
  const int chunk = lane % CPP;
  constexpr int UN = CPP >= 4 ? 4 : CPP;
  constexpr int NW = RawVec<T, VEC>::NW;
#pragma unroll 1
  for (int pass = 0; pass < (DUAL ? 2 : 1); ++pass) {
  if (DUAL && pass == 1) { xn = x2 + (size_t)n * xs2_n; on = out2 + (size_t)n * os2_n; }
#pragma unroll 1
  for (int sub0 = 0; sub0 < CPP; sub0 += UN) {
    uint32_t raw[UN][4][NW];
    float a[UN][4];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int src = (sub0 + u) * PPI + lane / CPP;
      const uint32_t b = __shfl_sync(0xffffffffu, base, src) + chunk * VEC;
      const uint32_t sxo = __shfl_sync(0xffffffffu, stepx, src);
      const uint32_t syo = __shfl_sync(0xffffffffu, stepy, src);
      a[u][0] = __shfl_sync(0xffffffffu, w00, src);
      a[u][1] = __shfl_sync(0xffffffffu, w01, src);
      a[u][2] = __shfl_sync(0xffffffffu, w10, src);
      a[u][3] = __shfl_sync(0xffffffffu, w11, src);
      RawVec<T, VEC>::ld(xn + b, raw[u][0]);              // dead pixels read element 0 with weight 0
      RawVec<T, VEC>::ld(xn + (uint32_t)(b + sxo), raw[u][1]);
      RawVec<T, VEC>::ld(xn + (uint32_t)(b + syo), raw[u][2]);
      RawVec<T, VEC>::ld(xn + (uint32_t)(b + syo + sxo), raw[u][3]);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int src = (sub0 + u) * PPI + lane / CPP;
      float r[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        r[j] = a[u][0] * RawVec<T, VEC>::get(raw[u][0], j) + a[u][1] * RawVec<T, VEC>::get(raw[u][1], j) +
               a[u][2] * RawVec<T, VEC>::get(raw[u][2], j) + a[u][3] * RawVec<T, VEC>::get(raw[u][3], j);
      const int oy = py0 + (src >> 3), ox = px0 + (src & 7);
      if ((live_mask >> src) & 1u) VecLoad<T, VEC>::st(on + (uint32_t)(oy * W + ox) * C + chunk * VEC, r);
    }
  }
  }
}

// ------------------------------------------------------------------------------------------
// Any-CPP vector path: NHWC, C * sizeof(T) multiple of 16 (e.g. 96 or 196 channels).
// ------------------------------------------------------------------------------------------
template <typename T, int PAD>
__global__ void __launch_bounds__(WARP_THREADS)
flow_warp_fwd_nhwc(const T* __restrict__ x, const float* __restrict__ flow, T* __restrict__ out, int C, int H,
                   int W, int layout, int tiles_x, int tiles_y, long long xs_n, long long os_n) {
  constexpr int VEC = 16 / sizeof(T);
  const int cpp = C / VEC;
  int tile = blockIdx.x;
  const int n = tile / (tiles_x * tiles_y);
  tile -= n * tiles_x * tiles_y;
  const int ty0 = (tile / tiles_x) * TILE_H, tx0 = (tile % tiles_x) * TILE_W;
  const T* xn = x + (size_t)n * xs_n;
  T* on = out + (size_t)n * os_n;
  const int items = TILE_PIX * cpp;
#pragma unroll 2
  for (int idx = threadIdx.x; idx < items; idx += WARP_THREADS) {
    const int p = idx / cpp, ch = idx - p * cpp;
    const int y = ty0 + p / TILE_W, xq = tx0 + p % TILE_W;
    if (y >= H || xq >= W) continue;
    float fx, fy;
    load_flow(flow, layout, n, y, xq, H, W, fx, fy);
    Corner c = make_corner<PAD>((float)y + fy, (float)xq + fx, H, W);
    float v[4][VEC];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (c.ok[k]) {
        VecLoad<T, VEC>::ld(xn + (size_t)c.off[k] * C + ch * VEC, v[k]);
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[k][j] = 0.f;
      }
    }
    float r[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j)
      r[j] = c.wgt[0] * v[0][j] + c.wgt[1] * v[1][j] + c.wgt[2] * v[2][j] + c.wgt[3] * v[3][j];
    VecLoad<T, VEC>::st(on + ((size_t)y * W + xq) * C + ch * VEC, r);
  }
}

// Backward of the fast path.  gx32 (fp32, NHWC, zero-filled by the host wrapper) receives the
// scatter; gflow is reduced over the channel chunks of a pixel with warp shuffles when the
// chunks of one pixel sit in one warp, otherwise with atomics.
template <typename T, int PAD, bool SHFL>
__global__ void __launch_bounds__(WARP_THREADS)
flow_warp_bwd_nhwc(const T* __restrict__ gout, const T* __restrict__ x, const float* __restrict__ flow,
                   float* __restrict__ gx32, float* __restrict__ gflow, int C, int H, int W, int layout,
                   int tiles_x, int tiles_y, long long gs_n, long long xs_n, long long gxs_n) {
  constexpr int VEC = 16 / sizeof(T);
  const int cpp = C / VEC;
  int tile = blockIdx.x;
  const int n = tile / (tiles_x * tiles_y);
  tile -= n * tiles_x * tiles_y;
  const int ty0 = (tile / tiles_x) * TILE_H, tx0 = (tile % tiles_x) * TILE_W;
  const T* xn = x + (size_t)n * xs_n;
  const T* gn = gout + (size_t)n * gs_n;
  float* gxn = gx32 ? gx32 + (size_t)n * gxs_n : nullptr;
  const int items = TILE_PIX * cpp;
  // SHFL: items % 256 == 0 and cpp | 32, so every warp iteration is full and pixel-aligned.
  for (int idx = threadIdx.x; idx < items; idx += WARP_THREADS) {
    const int p = idx / cpp, ch = idx - p * cpp;
    const int y = ty0 + p / TILE_W, xq = tx0 + p % TILE_W;
    const bool live = (y < H) && (xq < W);
    float dfx = 0.f, dfy = 0.f;
    if (live) {
      float fx, fy;
      load_flow(flow, layout, n, y, xq, H, W, fx, fy);
      Corner c = make_corner<PAD>((float)y + fy, (float)xq + fx, H, W);
      float g[VEC];
      VecLoad<T, VEC>::ld(gn + ((size_t)y * W + xq) * C + ch * VEC, g);
      if (gxn) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (c.ok[k]) {
            float* dst = gxn + (size_t)c.off[k] * C + ch * VEC;
#pragma unroll
            for (int j = 0; j < VEC; j += 4) {
              atomicAdd(reinterpret_cast<float4*>(dst + j),
                        make_float4(c.wgt[k] * g[j], c.wgt[k] * g[j + 1], c.wgt[k] * g[j + 2],
                                    c.wgt[k] * g[j + 3]));
            }
          }
        }
      }
      if (gflow) {
        float v[4][VEC];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (c.ok[k]) {
            VecLoad<T, VEC>::ld(xn + (size_t)c.off[k] * C + ch * VEC, v[k]);
          } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[k][j] = 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          dfx += g[j] * ((v[1][j] - v[0][j]) * (1.f - c.ly) + (v[3][j] - v[2][j]) * c.ly);
          dfy += g[j] * ((v[2][j] - v[0][j]) * (1.f - c.lx) + (v[3][j] - v[1][j]) * c.lx);
        }
        dfx *= (W == 1) ? 0.f : c.gxm;
        dfy *= (H == 1) ? 0.f : c.gym;
      }
    }
    if (gflow) {
      if (SHFL) {
        for (int s = cpp >> 1; s > 0; s >>= 1) {
          dfx += __shfl_xor_sync(0xffffffffu, dfx, s);
          dfy += __shfl_xor_sync(0xffffffffu, dfy, s);
        }
        if (live && ch == 0) {
          if (layout == EAVSR_FLOW_N2HW) {
            size_t b = ((size_t)n * 2) * H * W + (size_t)y * W + xq;
            gflow[b] = dfx;
            gflow[b + (size_t)H * W] = dfy;
          } else {
            reinterpret_cast<float2*>(gflow)[((size_t)n * H + y) * W + xq] = make_float2(dfx, dfy);
          }
        }
      } else if (live) {
        if (layout == EAVSR_FLOW_N2HW) {
          size_t b = ((size_t)n * 2) * H * W + (size_t)y * W + xq;
          atomicAdd(gflow + b, dfx);
          atomicAdd(gflow + b + (size_t)H * W, dfy);
        } else {
          size_t b = (((size_t)n * H + y) * W + xq) * 2;
          atomicAdd(gflow + b, dfx);
          atomicAdd(gflow + b + 1, dfy);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Strided path: any layout / any C (the 2-channel flow-composition warp and the 3-channel
// SPyNet image warps are NCHW with C*sizeof(T) < 16).  One thread per pixel, channel loop.
// ------------------------------------------------------------------------------------------
template <typename T, int PAD>
__global__ void __launch_bounds__(256)
flow_warp_fwd_strided(const T* __restrict__ x, Strides4 xs, const float* __restrict__ flow, T* __restrict__ out,
                      Strides4 os, int N, int C, int H, int W, int layout) {
  long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)N * H * W) return;
  int xq = (int)(pix % W);
  int y = (int)((pix / W) % H);
  int n = (int)(pix / ((long long)W * H));
  float fx, fy;
  load_flow(flow, layout, n, y, xq, H, W, fx, fy);
  Corner c = make_corner<PAD>((float)y + fy, (float)xq + fx, H, W);
  const T* xn = x + n * xs.n;
  long long o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = (long long)(c.off[k] / W) * xs.h + (long long)(c.off[k] % W) * xs.w;
  T* op = out + n * os.n + y * os.h + xq * os.w;
  for (int ch = 0; ch < C; ++ch) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (c.ok[k]) r += c.wgt[k] * to_f32<T>(xn[o[k] + ch * xs.c]);
    op[ch * os.c] = from_f32<T>(r);
  }
}

template <typename T, int PAD>
__global__ void __launch_bounds__(256)
flow_warp_bwd_strided(const T* __restrict__ gout, Strides4 gs, const T* __restrict__ x, Strides4 xs,
                      const float* __restrict__ flow, float* __restrict__ gx32, Strides4 gxs,
                      float* __restrict__ gflow, int N, int C, int H, int W, int layout) {
  long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)N * H * W) return;
  int xq = (int)(pix % W);
  int y = (int)((pix / W) % H);
  int n = (int)(pix / ((long long)W * H));
  float fx, fy;
  load_flow(flow, layout, n, y, xq, H, W, fx, fy);
  Corner c = make_corner<PAD>((float)y + fy, (float)xq + fx, H, W);
  const T* xn = x + n * xs.n;
  const T* gp = gout + n * gs.n + y * gs.h + xq * gs.w;
  float dfx = 0.f, dfy = 0.f;
  for (int ch = 0; ch < C; ++ch) {
    float g = to_f32<T>(gp[ch * gs.c]);
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int cy = c.off[k] / W, cx = c.off[k] % W;
      v[k] = c.ok[k] ? to_f32<T>(xn[(long long)cy * xs.h + (long long)cx * xs.w + ch * xs.c]) : 0.f;
      if (gx32 && c.ok[k])
        atomicAdd(gx32 + n * gxs.n + (long long)cy * gxs.h + (long long)cx * gxs.w + ch * gxs.c, c.wgt[k] * g);
    }
    dfx += g * ((v[1] - v[0]) * (1.f - c.ly) + (v[3] - v[2]) * c.ly);
    dfy += g * ((v[2] - v[0]) * (1.f - c.lx) + (v[3] - v[1]) * c.lx);
  }
  if (gflow) {
    dfx *= (W == 1) ? 0.f : c.gxm;
    dfy *= (H == 1) ? 0.f : c.gym;
    if (layout == EAVSR_FLOW_N2HW) {
      size_t b = ((size_t)n * 2) * H * W + (size_t)y * W + xq;
      gflow[b] = dfx;
      gflow[b + (size_t)H * W] = dfy;
    } else {
      reinterpret_cast<float2*>(gflow)[((size_t)n * H + y) * W + xq] = make_float2(dfx, dfy);
    }
  }
}

// ------------------------------------------------------------------------------------------
// backwarp: BaseModel.backwarp / get_backwarp (models/base_model.py:321-354) and
// PWCNET.Decoder.backwarp (models/pwc_net.py:184-207).  The reference appends a ones channel,
// calls grid_sample(align_corners=False, zeros) on a cached grid + flow / ((size-1)/2) and thresholds
// the warped ones (> 0.999) into a validity mask that multiplies the result.  Un-normalising that grid
// gives the sample point (y + fy*H/(H-1), x + fx*W/(W-1)); the warped ones channel is the sum of the
// in-image bilinear weights, so no ones channel, no grid and no cat ever exist here.
// One thread per pixel, channel loop over arbitrary strides (the callers are NCHW fp32: 3-channel HR
// frames and PWC pyramid features).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
backwarp_fwd_kernel(const T* __restrict__ x, Strides4 xs, const float* __restrict__ flow, T* __restrict__ out,
                    Strides4 os, T* __restrict__ mask, int N, int C, int H, int W, float ky, float kx) {
  long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)N * H * W) return;
  int xq = (int)(pix % W);
  int y = (int)((pix / W) % H);
  int n = (int)(pix / ((long long)W * H));
  float fx, fy;
  load_flow(flow, EAVSR_FLOW_N2HW, n, y, xq, H, W, fx, fy);
  Corner c = make_corner<EAVSR_PAD_ZEROS>((float)y + fy * ky, (float)xq + fx * kx, H, W);
  const float ones = c.wgt[0] + c.wgt[1] + c.wgt[2] + c.wgt[3];
  const float m = ones > 0.999f ? 1.f : 0.f;
  if (mask) mask[pix] = from_f32<T>(m);
  const T* xn = x + n * xs.n;
  long long o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = (long long)(c.off[k] / W) * xs.h + (long long)(c.off[k] % W) * xs.w;
  T* op = out + n * os.n + y * os.h + xq * os.w;
  for (int ch = 0; ch < C; ++ch) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (c.ok[k]) r += c.wgt[k] * to_f32<T>(xn[o[k] + ch * xs.c]);
    op[ch * os.c] = from_f32<T>(r * m);
  }
}

// Gradient of out = warp(x) * mask wrt x (scatter into the zero-filled fp32 gx32) and wrt flow.  The mask
// is piecewise constant (the reference overwrites it in place with 0 / 1), so it carries no gradient.
template <typename T>
__global__ void __launch_bounds__(256)
backwarp_bwd_kernel(const T* __restrict__ gout, Strides4 gs, const T* __restrict__ x, Strides4 xs,
                    const float* __restrict__ flow, float* __restrict__ gx32, Strides4 gxs,
                    float* __restrict__ gflow, int N, int C, int H, int W, float ky, float kx) {
  long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)N * H * W) return;
  int xq = (int)(pix % W);
  int y = (int)((pix / W) % H);
  int n = (int)(pix / ((long long)W * H));
  float fx, fy;
  load_flow(flow, EAVSR_FLOW_N2HW, n, y, xq, H, W, fx, fy);
  Corner c = make_corner<EAVSR_PAD_ZEROS>((float)y + fy * ky, (float)xq + fx * kx, H, W);
  const float ones = c.wgt[0] + c.wgt[1] + c.wgt[2] + c.wgt[3];
  const float m = ones > 0.999f ? 1.f : 0.f;
  const T* xn = x + n * xs.n;
  const T* gp = gout + n * gs.n + y * gs.h + xq * gs.w;
  float dfx = 0.f, dfy = 0.f;
  if (m != 0.f) {
    for (int ch = 0; ch < C; ++ch) {
      float g = to_f32<T>(gp[ch * gs.c]);
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int cy = c.off[k] / W, cx = c.off[k] % W;
        v[k] = c.ok[k] ? to_f32<T>(xn[(long long)cy * xs.h + (long long)cx * xs.w + ch * xs.c]) : 0.f;
        if (gx32 && c.ok[k])
          atomicAdd(gx32 + n * gxs.n + (long long)cy * gxs.h + (long long)cx * gxs.w + ch * gxs.c, c.wgt[k] * g);
      }
      dfx += g * ((v[1] - v[0]) * (1.f - c.ly) + (v[3] - v[2]) * c.ly);
      dfy += g * ((v[2] - v[0]) * (1.f - c.lx) + (v[3] - v[1]) * c.lx);
    }
  }
  if (gflow) {
    size_t b = ((size_t)n * 2) * H * W + (size_t)y * W + xq;
    gflow[b] = dfx * kx;
    gflow[b + (size_t)H * W] = dfy * ky;
  }
}

// ------------------------------------------------------------------------------------------
// SPyNet level input (SPyNet.compute_flow, models/eavsrp_model.py:468-486): per pyramid level
// the reference runs F.interpolate(flow, 2) * 2, a permute, flow_warp(supp, ., 'border') and torch.cat([ref, warped,
// flow_up]) -- five launches around three 3-channel images.  One kernel writes the 8-channel input of the level's
// first 7x7 convolution: channels 0-2 ref, 3-5 supp warped (border padding) by up = 2 * resize_x2(flow_prev), 6-7 up.
// NCHW fp32 (flows are coordinates: SPyNet stays fp32), one thread per pixel; flow_prev == NULL: up = 0 (level 0).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
spynet_level_input_kernel(const float* __restrict__ ref, const float* __restrict__ supp,
                          const float* __restrict__ flow_prev, float* __restrict__ out, int N, int H, int W, int hp,
                          int wp, float ry, float rx) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)N * H * W) return;
  const int x = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / ((long long)W * H));
  const size_t plane = (size_t)H * W, o = (size_t)y * W + x;
  float ux = 0.f, uy = 0.f;
  if (flow_prev) {
    const float sy = ry * (float)y, sx = rx * (float)x;
    const int y0 = (int)sy, x0 = (int)sx;
    const int yp = (y0 < hp - 1) ? wp : 0, xp = (x0 < wp - 1) ? 1 : 0;
    const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* p = flow_prev + (size_t)n * 2 * hp * wp + (size_t)y0 * wp + x0;
    ux = (hy * (hx * __ldg(p) + lx * __ldg(p + xp)) + ly * (hx * __ldg(p + yp) + lx * __ldg(p + yp + xp))) * 2.f;
    p += (size_t)hp * wp;
    uy = (hy * (hx * __ldg(p) + lx * __ldg(p + xp)) + ly * (hx * __ldg(p + yp) + lx * __ldg(p + yp + xp))) * 2.f;
  }
  const Corner c = make_corner<EAVSR_PAD_BORDER>((float)y + (H == 1 ? 0.f : uy), (float)x + (W == 1 ? 0.f : ux), H, W);
  float* on = out + (size_t)n * 8 * plane + o;
  const float* rn = ref + (size_t)n * 3 * plane + o;
  const float* sn = supp + (size_t)n * 3 * plane;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    on[ch * plane] = __ldg(rn + ch * plane);
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (c.ok[k]) r += c.wgt[k] * __ldg(sn + ch * plane + c.off[k]);
    on[(3 + ch) * plane] = r;
  }
  on[6 * plane] = ux;
  on[7 * plane] = uy;
}

bool is_nhwc_dense(const int64_t s[4], int c, int h, int w) {
  return s[1] == 1 && s[3] == c && s[2] == (int64_t)w * c && s[0] >= (int64_t)h * w * c;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
int fwd_dispatch(const void* x, const int64_t* xs, const float* flow, int layout, void* out, const int64_t* os,
                 int n, int c, int h, int w, int pad, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  const bool fast = is_nhwc_dense(xs, c, h, w) && is_nhwc_dense(os, c, h, w) && (c % VEC == 0) && aligned16(x) &&
                    aligned16(out) && ((xs[0] * sizeof(T)) % 16 == 0) && ((os[0] * sizeof(T)) % 16 == 0);
  const int cpp = c / VEC;
  const bool lean = fast && cpp <= 32 && (cpp & (cpp - 1)) == 0 && ((long long)h * w * c < (1ll << 31));
  if (lean) {
    int tx = ceil_div(w, TILE_W), ty = ceil_div(h, TILE_H);
    long long blocks = (long long)n * tx * ty;
#define EAVSR_LEAN(CPPV)                                                                                   \
  case CPPV:                                                                                               \
    if (pad == EAVSR_PAD_ZEROS)                                                                            \
      flow_warp_fwd_lean<T, CPPV, EAVSR_PAD_ZEROS><<<(unsigned)blocks, LEAN_THREADS, 0, st>>>(             \
          (const T*)x, flow, (T*)out, h, w, layout, tx, ty, xs[0], os[0]);                                 \
    else                                                                                                   \
      flow_warp_fwd_lean<T, CPPV, EAVSR_PAD_BORDER><<<(unsigned)blocks, LEAN_THREADS, 0, st>>>(            \
          (const T*)x, flow, (T*)out, h, w, layout, tx, ty, xs[0], os[0]);                                 \
    break;
    switch (cpp) {
      EAVSR_LEAN(1) EAVSR_LEAN(2) EAVSR_LEAN(4) EAVSR_LEAN(8) EAVSR_LEAN(16) EAVSR_LEAN(32)
    }
#undef EAVSR_LEAN
  } else if (fast) {
    int tx = ceil_div(w, TILE_W), ty = ceil_div(h, TILE_H);
    long long blocks = (long long)n * tx * ty;
    if (pad == EAVSR_PAD_ZEROS)
      flow_warp_fwd_nhwc<T, EAVSR_PAD_ZEROS><<<(unsigned)blocks, WARP_THREADS, 0, st>>>(
          (const T*)x, flow, (T*)out, c, h, w, layout, tx, ty, xs[0], os[0]);
    else
      flow_warp_fwd_nhwc<T, EAVSR_PAD_BORDER><<<(unsigned)blocks, WARP_THREADS, 0, st>>>(
          (const T*)x, flow, (T*)out, c, h, w, layout, tx, ty, xs[0], os[0]);
  } else {
    Strides4 a{xs[0], xs[1], xs[2], xs[3]}, b{os[0], os[1], os[2], os[3]};
    long long blocks = ((long long)n * h * w + 255) / 256;
    if (pad == EAVSR_PAD_ZEROS)
      flow_warp_fwd_strided<T, EAVSR_PAD_ZEROS><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, a, flow, (T*)out, b,
                                                                                  n, c, h, w, layout);
    else
      flow_warp_fwd_strided<T, EAVSR_PAD_BORDER><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, a, flow, (T*)out,
                                                                                   b, n, c, h, w, layout);
  }
  return check_launch("flow_warp_forward");
}

template <typename T>
int bwd_dispatch(const void* gout, const int64_t* gs, const void* x, const int64_t* xs, const float* flow,
                 int layout, float* gx32, const int64_t* gxs, float* gflow, int n, int c, int h, int w, int pad,
                 cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  const int cpp = c / VEC;
  const bool fast = is_nhwc_dense(xs, c, h, w) && is_nhwc_dense(gs, c, h, w) &&
                    (!gx32 || is_nhwc_dense(gxs, c, h, w)) && (c % VEC == 0) && aligned16(x) && aligned16(gout) &&
                    (!gx32 || aligned16(gx32)) && ((xs[0] * sizeof(T)) % 16 == 0) &&
                    ((gs[0] * sizeof(T)) % 16 == 0) && (!gx32 || (gxs[0] * 4) % 16 == 0);
  if (fast) {
    const bool shfl = (cpp >= 2) && (cpp <= 32) && ((cpp & (cpp - 1)) == 0);
    int tx = ceil_div(w, TILE_W), ty = ceil_div(h, TILE_H);
    long long blocks = (long long)n * tx * ty;
    if (gflow && !shfl && cpp > 1) {
      cudaMemsetAsync(gflow, 0, (size_t)n * 2 * h * w * sizeof(float), st);
    }
#define EAVSR_LAUNCH_BWD(PADV, SH)                                                                       \
  flow_warp_bwd_nhwc<T, PADV, SH><<<(unsigned)blocks, WARP_THREADS, 0, st>>>(                            \
      (const T*)gout, (const T*)x, flow, gx32, gflow, c, h, w, layout, tx, ty, gs[0], xs[0], gx32 ? gxs[0] : 0)
    // cpp == 1: one thread owns the pixel; the SHFL variant degenerates to a plain store.
    const bool sh = shfl || cpp == 1;
    if (pad == EAVSR_PAD_ZEROS) { if (sh) EAVSR_LAUNCH_BWD(EAVSR_PAD_ZEROS, true); else EAVSR_LAUNCH_BWD(EAVSR_PAD_ZEROS, false); }
    else                        { if (sh) EAVSR_LAUNCH_BWD(EAVSR_PAD_BORDER, true); else EAVSR_LAUNCH_BWD(EAVSR_PAD_BORDER, false); }
#undef EAVSR_LAUNCH_BWD
  } else {
    Strides4 a{gs[0], gs[1], gs[2], gs[3]}, b{xs[0], xs[1], xs[2], xs[3]}, g{0, 0, 0, 0};
    if (gx32) g = Strides4{gxs[0], gxs[1], gxs[2], gxs[3]};
    long long blocks = ((long long)n * h * w + 255) / 256;
    if (pad == EAVSR_PAD_ZEROS)
      flow_warp_bwd_strided<T, EAVSR_PAD_ZEROS><<<(unsigned)blocks, 256, 0, st>>>(
          (const T*)gout, a, (const T*)x, b, flow, gx32, g, gflow, n, c, h, w, layout);
    else
      flow_warp_bwd_strided<T, EAVSR_PAD_BORDER><<<(unsigned)blocks, 256, 0, st>>>(
          (const T*)gout, a, (const T*)x, b, flow, gx32, g, gflow, n, c, h, w, layout);
  }
  return check_launch("flow_warp_backward");
}

// extent (in elements) of a strided (n,c,h,w) tensor, for zero-filling accumulation buffers
size_t strided_extent(const int64_t s[4], int n, int c, int h, int w) {
  return (size_t)((n - 1) * s[0] + (c - 1) * s[1] + (h - 1) * s[2] + (w - 1) * s[3] + 1);
}

}  // namespace

size_t strided_extent_elems(const int64_t s[4], int n, int c, int h, int w) { return strided_extent(s, n, c, h, w); }

}  // namespace eavsr

using namespace eavsr;

extern "C" int eavsr_flow_warp_forward(const void* x, const int64_t x_strides[4], const float* flow,
                                       int flow_layout, void* out, const int64_t out_strides[4], int n, int c,
                                       int h, int w, int dtype, int padding_mode, void* stream) {
  EAVSR_REQUIRE(x && flow && out && x_strides && out_strides, "flow_warp_forward: null pointer");
  EAVSR_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "flow_warp_forward: empty tensor (n=%d c=%d h=%d w=%d)", n, c, h, w);
  EAVSR_REQUIRE(flow_layout == EAVSR_FLOW_N2HW || flow_layout == EAVSR_FLOW_NHW2, "flow_warp_forward: bad flow_layout %d", flow_layout);
  EAVSR_REQUIRE(padding_mode == EAVSR_PAD_ZEROS || padding_mode == EAVSR_PAD_BORDER,
                "flow_warp_forward: padding_mode %d not supported (zeros/border only)", padding_mode);
  EAVSR_REQUIRE((long long)h * w < (1ll << 31), "flow_warp_forward: image too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == EAVSR_F32) return fwd_dispatch<float>(x, x_strides, flow, flow_layout, out, out_strides, n, c, h, w, padding_mode, st);
  if (dtype == EAVSR_BF16) return fwd_dispatch<__nv_bfloat16>(x, x_strides, flow, flow_layout, out, out_strides, n, c, h, w, padding_mode, st);
  set_error("flow_warp_forward: bad dtype %d", dtype);
  return EAVSR_ERR_INVALID;
}

extern "C" int eavsr_flow_warp_backward(const void* gout, const int64_t gout_strides[4], const void* x,
                                        const int64_t x_strides[4], const float* flow, int flow_layout, float* gx32,
                                        const int64_t gx_strides[4], float* gflow, int n, int c, int h, int w,
                                        int dtype, int padding_mode, void* stream) {
  EAVSR_REQUIRE(gout && x && flow && gout_strides && x_strides, "flow_warp_backward: null pointer");
  EAVSR_REQUIRE(!gx32 || gx_strides, "flow_warp_backward: gx32 without strides");
  EAVSR_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "flow_warp_backward: empty tensor");
  EAVSR_REQUIRE(flow_layout == EAVSR_FLOW_N2HW || flow_layout == EAVSR_FLOW_NHW2, "flow_warp_backward: bad flow_layout %d", flow_layout);
  EAVSR_REQUIRE(padding_mode == EAVSR_PAD_ZEROS || padding_mode == EAVSR_PAD_BORDER, "flow_warp_backward: bad padding_mode %d", padding_mode);
  cudaStream_t st = (cudaStream_t)stream;
  if (!gx32 && !gflow) return EAVSR_OK;
  if (gx32) {
    cudaError_t e = cudaMemsetAsync(gx32, 0, strided_extent(gx_strides, n, c, h, w) * sizeof(float), st);
    if (e != cudaSuccess) { set_error("flow_warp_backward: memset: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
  }
  if (dtype == EAVSR_F32) return bwd_dispatch<float>(gout, gout_strides, x, x_strides, flow, flow_layout, gx32, gx_strides, gflow, n, c, h, w, padding_mode, st);
  if (dtype == EAVSR_BF16) return bwd_dispatch<__nv_bfloat16>(gout, gout_strides, x, x_strides, flow, flow_layout, gx32, gx_strides, gflow, n, c, h, w, padding_mode, st);
  set_error("flow_warp_backward: bad dtype %d", dtype);
  return EAVSR_ERR_INVALID;
}

extern "C" int eavsr_flow_warp2_forward(const void* x1, const int64_t x1_strides[4], const void* x2,
                                        const int64_t x2_strides[4], const float* flow, int flow_layout, void* out1,
                                        const int64_t out1_strides[4], void* out2, const int64_t out2_strides[4],
                                        int n, int c, int h, int w, int dtype, int padding_mode, void* stream) {
  EAVSR_REQUIRE(x1 && x2 && flow && out1 && out2 && x1_strides && x2_strides && out1_strides && out2_strides,
                "flow_warp2_forward: null pointer");
  EAVSR_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "flow_warp2_forward: empty tensor");
  EAVSR_REQUIRE(flow_layout == EAVSR_FLOW_N2HW || flow_layout == EAVSR_FLOW_NHW2, "flow_warp2_forward: bad flow_layout %d", flow_layout);
  EAVSR_REQUIRE(padding_mode == EAVSR_PAD_ZEROS || padding_mode == EAVSR_PAD_BORDER, "flow_warp2_forward: bad padding_mode %d", padding_mode);
  auto ok = [&](const void* p, const int64_t* s) {
    return is_nhwc_dense(s, c, h, w) && aligned16(p) && (s[0] * 2) % 16 == 0;
  };
  if (dtype != EAVSR_BF16 || c != 64 || (long long)h * w * c >= (1ll << 31) || !ok(x1, x1_strides) ||
      !ok(x2, x2_strides) || !ok(out1, out1_strides) || !ok(out2, out2_strides)) {
    set_error("flow_warp2_forward: only dense NHWC bf16 64-channel maps are fused (call flow_warp_forward twice)");
    return EAVSR_ERR_UNSUPPORTED;
  }
  using T = __nv_bfloat16;
  const int tx = ceil_div(w, TILE_W), ty = ceil_div(h, TILE_H);
  const long long blocks = (long long)n * tx * ty;
  cudaStream_t st = (cudaStream_t)stream;
  if (padding_mode == EAVSR_PAD_ZEROS)
    flow_warp_fwd_lean<T, 8, EAVSR_PAD_ZEROS, true><<<(unsigned)blocks, LEAN_THREADS, 0, st>>>(
        (const T*)x1, flow, (T*)out1, h, w, flow_layout, tx, ty, x1_strides[0], out1_strides[0], (const T*)x2, (T*)out2,
        x2_strides[0], out2_strides[0]);
  else
    flow_warp_fwd_lean<T, 8, EAVSR_PAD_BORDER, true><<<(unsigned)blocks, LEAN_THREADS, 0, st>>>(
        (const T*)x1, flow, (T*)out1, h, w, flow_layout, tx, ty, x1_strides[0], out1_strides[0], (const T*)x2, (T*)out2,
        x2_strides[0], out2_strides[0]);
  return check_launch("flow_warp2_forward");
}

extern "C" int eavsr_backwarp_forward(const void* x, const int64_t x_strides[4], const float* flow, void* out,
                                      const int64_t out_strides[4], void* mask, int n, int c, int h, int w,
                                      int dtype, void* stream) {
  EAVSR_REQUIRE(x && flow && out && x_strides && out_strides, "backwarp_forward: null pointer");
  EAVSR_REQUIRE(n > 0 && c > 0, "backwarp_forward: empty tensor (n=%d c=%d)", n, c);
  EAVSR_REQUIRE(h > 1 && w > 1, "backwarp_forward: h and w must be > 1 (the reference divides by (size-1)/2), got %dx%d", h, w);
  EAVSR_REQUIRE((long long)h * w < (1ll << 31), "backwarp_forward: image too large");
  cudaStream_t st = (cudaStream_t)stream;
  const float ky = (float)h / (float)(h - 1), kx = (float)w / (float)(w - 1);
  Strides4 a{x_strides[0], x_strides[1], x_strides[2], x_strides[3]};
  Strides4 b{out_strides[0], out_strides[1], out_strides[2], out_strides[3]};
  const long long blocks = ((long long)n * h * w + 255) / 256;
  if (dtype == EAVSR_F32)
    backwarp_fwd_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)x, a, flow, (float*)out, b, (float*)mask,
                                                                 n, c, h, w, ky, kx);
  else if (dtype == EAVSR_BF16)
    backwarp_fwd_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(
        (const __nv_bfloat16*)x, a, flow, (__nv_bfloat16*)out, b, (__nv_bfloat16*)mask, n, c, h, w, ky, kx);
  else { set_error("backwarp_forward: bad dtype %d", dtype); return EAVSR_ERR_INVALID; }
  return check_launch("backwarp_forward");
}

extern "C" int eavsr_backwarp_backward(const void* gout, const int64_t gout_strides[4], const void* x,
                                       const int64_t x_strides[4], const float* flow, float* gx32,
                                       const int64_t gx_strides[4], float* gflow, int n, int c, int h, int w,
                                       int dtype, void* stream) {
  EAVSR_REQUIRE(gout && x && flow && gout_strides && x_strides, "backwarp_backward: null pointer");
  EAVSR_REQUIRE(!gx32 || gx_strides, "backwarp_backward: gx32 without strides");
  EAVSR_REQUIRE(n > 0 && c > 0 && h > 1 && w > 1, "backwarp_backward: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (!gx32 && !gflow) return EAVSR_OK;
  if (gx32) {
    cudaError_t e = cudaMemsetAsync(gx32, 0, strided_extent(gx_strides, n, c, h, w) * sizeof(float), st);
    if (e != cudaSuccess) { set_error("backwarp_backward: memset: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
  }
  const float ky = (float)h / (float)(h - 1), kx = (float)w / (float)(w - 1);
  Strides4 a{gout_strides[0], gout_strides[1], gout_strides[2], gout_strides[3]};
  Strides4 b{x_strides[0], x_strides[1], x_strides[2], x_strides[3]}, g{0, 0, 0, 0};
  if (gx32) g = Strides4{gx_strides[0], gx_strides[1], gx_strides[2], gx_strides[3]};
  const long long blocks = ((long long)n * h * w + 255) / 256;
  if (dtype == EAVSR_F32)
    backwarp_bwd_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)gout, a, (const float*)x, b, flow, gx32,
                                                                 g, gflow, n, c, h, w, ky, kx);
  else if (dtype == EAVSR_BF16)
    backwarp_bwd_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(
        (const __nv_bfloat16*)gout, a, (const __nv_bfloat16*)x, b, flow, gx32, g, gflow, n, c, h, w, ky, kx);
  else { set_error("backwarp_backward: bad dtype %d", dtype); return EAVSR_ERR_INVALID; }
  return check_launch("backwarp_backward");
}

extern "C" int eavsr_flow_warp_pyramid_forward(const void* x, const int64_t x_strides[4], const void* x2,
                                               const int64_t x2_strides[4], const EavsrFlowTerm* terms, int nterms,
                                               void* out, const int64_t out_strides[4], void* out2,
                                               const int64_t out2_strides[4], float* flow_out, int n, int c, int h,
                                               int w, int dtype, int padding_mode, void* stream) {
  EAVSR_REQUIRE(x && terms && out && x_strides && out_strides, "flow_warp_pyramid_forward: null pointer");
  EAVSR_REQUIRE(!x2 || (out2 && x2_strides && out2_strides), "flow_warp_pyramid_forward: x2 without out2");
  EAVSR_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "flow_warp_pyramid_forward: empty tensor");
  EAVSR_REQUIRE(nterms >= 1 && nterms <= PYR_MAX, "flow_warp_pyramid_forward: 1..%d flow terms (got %d)", PYR_MAX, nterms);
  EAVSR_REQUIRE(padding_mode == EAVSR_PAD_ZEROS, "flow_warp_pyramid_forward: zeros padding only");
  PyrParams P{};
  P.nterms = nterms;
  P.flow_out = flow_out;
  for (int i = 0; i < nterms; ++i) {
    EAVSR_REQUIRE(terms[i].flow && terms[i].h > 0 && terms[i].w > 0, "flow_warp_pyramid_forward: bad term %d", i);
    P.t[i] = PyrTerm{terms[i].flow, terms[i].scaled_out, terms[i].h, terms[i].w, terms[i].scale,
                     h > 1 ? (float)(terms[i].h - 1) / (float)(h - 1) : 0.f, w > 1 ? (float)(terms[i].w - 1) / (float)(w - 1) : 0.f};
  }
  auto dense = [&](const void* p, const int64_t* s, size_t es) {
    return is_nhwc_dense(s, c, h, w) && aligned16(p) && (s[0] * es) % 16 == 0;
  };
  const size_t es = dtype == EAVSR_BF16 ? 2 : 4;
  const bool ok = (dtype == EAVSR_BF16 || dtype == EAVSR_F32) && c == 64 && (long long)h * w * c < (1ll << 31) &&
                  dense(x, x_strides, es) && dense(out, out_strides, es) &&
                  (!x2 || (dtype == EAVSR_BF16 && dense(x2, x2_strides, es) && dense(out2, out2_strides, es)));
  if (!ok) {
    set_error("flow_warp_pyramid_forward: dense NHWC 64-channel bf16 / fp32 maps only (dual: bf16); compose "
              "F.interpolate + flow_warp_forward otherwise");
    return EAVSR_ERR_UNSUPPORTED;
  }
  const int tx = ceil_div(w, TILE_W), ty = ceil_div(h, TILE_H);
  const unsigned blocks = (unsigned)((long long)n * tx * ty);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == EAVSR_BF16) {
    using T = __nv_bfloat16;
    if (x2)
      flow_warp_fwd_lean<T, 8, EAVSR_PAD_ZEROS, true, true><<<blocks, LEAN_THREADS, 0, st>>>(
          (const T*)x, nullptr, (T*)out, h, w, EAVSR_FLOW_N2HW, tx, ty, x_strides[0], out_strides[0], (const T*)x2,
          (T*)out2, x2_strides[0], out2_strides[0], P);
    else
      flow_warp_fwd_lean<T, 8, EAVSR_PAD_ZEROS, false, true><<<blocks, LEAN_THREADS, 0, st>>>(
          (const T*)x, nullptr, (T*)out, h, w, EAVSR_FLOW_N2HW, tx, ty, x_strides[0], out_strides[0], nullptr, nullptr,
          0, 0, P);
  } else {
    flow_warp_fwd_lean<float, 16, EAVSR_PAD_ZEROS, false, true><<<blocks, LEAN_THREADS, 0, st>>>(
        (const float*)x, nullptr, (float*)out, h, w, EAVSR_FLOW_N2HW, tx, ty, x_strides[0], out_strides[0], nullptr,
        nullptr, 0, 0, P);
  }
  return check_launch("flow_warp_pyramid_forward");
}

extern "C" int eavsr_spynet_level_input_forward(const float* ref, const float* supp, const float* flow_prev, float* out,
                                                int n, int h, int w, int prev_h, int prev_w, void* stream) {
  EAVSR_REQUIRE(ref && supp && out, "spynet_level_input_forward: null pointer");
  EAVSR_REQUIRE(n > 0 && h > 0 && w > 0, "spynet_level_input_forward: empty tensor");
  EAVSR_REQUIRE(!flow_prev || (prev_h > 0 && prev_w > 0), "spynet_level_input_forward: bad previous flow size");
  EAVSR_REQUIRE((long long)h * w < (1ll << 31), "spynet_level_input_forward: image too large");
  const long long total = (long long)n * h * w;
  const float ry = (flow_prev && h > 1) ? (float)(prev_h - 1) / (float)(h - 1) : 0.f;
  const float rx = (flow_prev && w > 1) ? (float)(prev_w - 1) / (float)(w - 1) : 0.f;
  spynet_level_input_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      ref, supp, flow_prev, out, n, h, w, prev_h, prev_w, ry, rx);
  return check_launch("spynet_level_input_forward");
}
