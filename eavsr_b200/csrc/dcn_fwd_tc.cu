// DCNv2 forward as an implicit GEMM on tcgen05 / TMEM (sm_100a), plus the C-ABI dispatch.
//
// Replaces mmcv's modulated_deformable_im2col (576 x P fp32 column buffer in HBM) + cuBLAS
// GEMM under models/networks.py:627-630 for the model's configuration: 64 -> 64 channels,
// 3x3, stride 1, pad 1, dilation 1, groups 1, deform_groups in {1,2,4,8,16}, NHWC features.
//
//   D[128 pixels x 64 cout] = sum_{tap=0..8}  A_tap[128 x 64 cin] * W_tap[64 cin x 64 cout]
//
// A_tap is never in HBM: the CTA's 256 threads bilinear-sample x at (y-1+i+dy, x-1+j+dx),
// multiply by the modulation mask and write bf16 straight into a 128-byte-swizzled K-major
// shared-memory tile that one thread hands to tcgen05.mma; the fp32 accumulator lives in TMEM
// and is read back once per tile (tcgen05.ld) for bias + store.  Offsets/masks of the next tap
// and the next tap's packed weights stream in with cp.async while the current tap is gathered.
// fp32 features use a bf16x3 split (hi*hi + lo*hi + hi*lo) so the result stays within ~1e-5 of
// fp32 math; bf16 features use one MMA per K-step.
//
// Work decomposition: tile = 128 consecutive pixels of one image, persistent CTAs
// (2 per SM when shared memory allows) striding over tiles.
#include "common.cuh"
#include "dcn_fwd_ws.cuh"
#include "dcn_fwd_win.cuh"
#include "dcn_fwd_win2.cuh"
#include "dcn_fwd_win3.cuh"

namespace eavsr {

template <typename T>
int dcn_forward_generic(const void* x, const int64_t* xs, const float* offset, const float* mask,
                        const void* weight, const void* bias, void* out, const int64_t* os, const DcnGeom& g,
                        cudaStream_t st);
template <typename T>
int dcn_backward_generic(const void* gout, const int64_t* gs, const void* x, const int64_t* xs,
                         const float* offset, const float* mask, const void* weight, float* gx32,
                         const int64_t* gxs, float* goffset, float* gmask, float* gweight32, float* gbias32,
                         const DcnGeom& g, cudaStream_t st);

namespace {

constexpr int TC_THREADS = 256;
constexpr int TILE_M = 128;           // pixels per tile == UMMA M
constexpr int CH = 64;                // cin == cout == UMMA N == K per tap
constexpr int TAPS = 9;
constexpr int A_TILE_BYTES = TILE_M * CH * 2;  // 16 KB bf16
constexpr int B_TILE_BYTES = CH * CH * 2;      // 8 KB bf16
constexpr int NB_STAGES = 3;
constexpr int NA_STAGES = 2;
constexpr int TMEM_COLS = 64;

template <int DG> struct TcCfg {
  static constexpr int LPP = DG <= 8 ? 8 : 16;     // lanes (threads) per pixel
  static constexpr int CPI = CH / LPP;             // channels per work item
  static constexpr int NI = TILE_M * LPP / TC_THREADS;  // items per thread per tap
  static constexpr int ROWS_PER_PASS = TC_THREADS / LPP;
  static constexpr int PL = DG <= 8 ? 132 : 130;   // padded plane length (bank-conflict free)
  static constexpr int NPLANES = 3 * DG;
  static constexpr int OFF_BYTES = NPLANES * PL * 4;
};

template <bool SPLIT, int DG> struct TcSmem {
  static constexpr int TERMS = SPLIT ? 2 : 1;
  static constexpr int B_STAGE = B_TILE_BYTES * TERMS;
  static constexpr int A_STAGE = A_TILE_BYTES * TERMS;
  static constexpr int OFF_STAGE = (TcCfg<DG>::OFF_BYTES + 15) / 16 * 16;
  static constexpr int B_OFF = 0;
  static constexpr int A_OFF = B_OFF + NB_STAGES * B_STAGE;
  static constexpr int OFFS_OFF = A_OFF + NA_STAGES * A_STAGE;
  static constexpr int BAR_OFF = OFFS_OFF + 2 * OFF_STAGE;
  static constexpr int TOTAL = BAR_OFF + 64;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024 B alignment
};

// Packed weights: [tap][term(hi,lo)][8 KB image of a 64x64 K-major SW128 tile], row = cout, k = cin.
template <typename T, bool SPLIT>
__global__ void dcn_pack_weight(const T* __restrict__ w, uint8_t* __restrict__ packed) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (tap, o, c)
  if (idx >= TAPS * CH * CH) return;
  const int c = idx % CH, o = (idx / CH) % CH, t = idx / (CH * CH);
  const float v = to_f32<T>(w[((size_t)o * CH + c) * TAPS + t]);
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  constexpr int TERMS = SPLIT ? 2 : 1;
  uint8_t* base = packed + (size_t)t * TERMS * B_TILE_BYTES;
  const uint32_t off = sw128_offset(o, c * 2);
  *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
  if (SPLIT) {
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    *reinterpret_cast<__nv_bfloat16*>(base + B_TILE_BYTES + off) = lo;
  }
}

template <typename XT, bool SPLIT, int DG, bool VEC_OFF>
__global__ void __launch_bounds__(TC_THREADS, SPLIT ? 1 : 2)
dcn_fwd_tc_kernel(const XT* __restrict__ x, const float* __restrict__ offset, const float* __restrict__ mask,
                  const uint8_t* __restrict__ wpacked, const XT* __restrict__ bias, XT* __restrict__ out,
                  int H, int W, long long xs_n, long long os_n, int tiles_per_img, int total_tiles) {
  using Cfg = TcCfg<DG>;
  using SM = TcSmem<SPLIT, DG>;
  constexpr int LPP = Cfg::LPP, CPI = Cfg::CPI, NI = Cfg::NI, PL = Cfg::PL;
  constexpr int TERMS = SM::TERMS;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sB = smem_base + SM::B_OFF, sA = smem_base + SM::A_OFF, sOff = smem_base + SM::OFFS_OFF;
  const float* sOffF = reinterpret_cast<const float*>(smem + SM::OFFS_OFF);
  const uint32_t bar0 = smem_base + SM::BAR_OFF;            // mma_done[0], [1] at +8
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM::BAR_OFF + 32);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int HW = H * W;

  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(smem_base + SM::BAR_OFF + 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  // asynchronous stage fill: offsets + mask planes of (tile, tap) and the tap's packed weights
  auto prefetch = [&](int tile, int tap, int off_buf, int b_stage) {
    const int n = tile / tiles_per_img;
    const int pix0 = (tile - n * tiles_per_img) * TILE_M;
    const uint32_t dstO = sOff + off_buf * SM::OFF_STAGE;
    constexpr int VW = VEC_OFF ? (DG <= 8 ? 4 : 2) : 1;  // floats per cp.async
    constexpr int PER_PLANE = TILE_M / VW;
    for (int i = tid; i < Cfg::NPLANES * PER_PLANE; i += TC_THREADS) {
      const int plane = i / PER_PLANE, r = (i - plane * PER_PLANE) * VW;
      const int comp = plane / DG, g = plane - comp * DG;
      if (pix0 + r < HW) {
        const float* src = (comp < 2)
            ? offset + ((size_t)(n * DG + g) * TAPS + tap) * 2 * HW + (size_t)comp * HW + pix0 + r
            : mask + ((size_t)(n * DG + g) * TAPS + tap) * HW + pix0 + r;
        const uint32_t dst = dstO + (plane * PL + r) * 4;
        if (VW == 4) cp_async_16(dst, src);
        else if (VW == 2) cp_async_8(dst, src);
        else cp_async_4(dst, src);
      }
    }
    const uint8_t* wsrc = wpacked + (size_t)tap * SM::B_STAGE;
    const uint32_t dstB = sB + b_stage * SM::B_STAGE;
    for (int i = tid; i < SM::B_STAGE / 16; i += TC_THREADS) cp_async_16(dstB + i * 16, wsrc + i * 16);
    cp_async_commit();
  };

  const int l = tid % LPP;                 // lane-in-pixel: channel chunk
  const int rbase = tid / LPP;
  const int grp = (l * DG) / LPP;          // deformable group of this chunk
  constexpr uint32_t IDESC = umma_idesc_bf16(TILE_M, CH);

  int it = 0;  // global (tile, tap) iteration counter of this CTA
  int tile = blockIdx.x;
  if (tile < total_tiles) prefetch(tile, 0, 0, 0);

  for (; tile < total_tiles; tile += gridDim.x) {
    const int n = tile / tiles_per_img;
    const int pix0 = (tile - n * tiles_per_img) * TILE_M;
    const XT* xn = x + (size_t)n * xs_n;
    float py_f[NI], px_f[NI];   // un-offset sampling position of tap (0,0): (y-1, x-1)
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int pix = pix0 + rbase + j * Cfg::ROWS_PER_PASS;
      const int yy = pix / W;
      py_f[j] = pix < HW ? (float)(yy - 1) : -100000.f;   // sentinel: sample rejected as out of range
      px_f[j] = pix < HW ? (float)(pix - yy * W - 1) : 0.f;
    }

    for (int tap = 0; tap < TAPS; ++tap, ++it) {
      const int s = it & 1;
      cp_async_wait<0>();
      __syncthreads();
      if (it >= 2) mbar_wait(bar0 + 8 * s, ((it >> 1) - 1) & 1);  // MMA(it-2) done: A[s], B[(it+1)%3] free
      {
        int ntile = tile, ntap = tap + 1;
        if (ntap == TAPS) { ntap = 0; ntile += gridDim.x; }
        if (ntile < total_tiles) prefetch(ntile, ntap, s ^ 1, (it + 1) % NB_STAGES);
      }
      const float* so = sOffF + (size_t)s * (SM::OFF_STAGE / 4);
      const uint32_t aStage = sA + s * SM::A_STAGE;
      const int ti = tap / 3, tj = tap - ti * 3;
      constexpr int JB = (sizeof(XT) == 4 && CPI == 8) ? 2 : 4;  // items gathered per batch
#pragma unroll
      for (int j0 = 0; j0 < NI; j0 += JB) {
        uint32_t raw[JB][4][RawVec<XT, CPI>::NW];
        float wgt[JB][4];
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
          const int j = j0 + jj;
          const int r = rbase + j * Cfg::ROWS_PER_PASS;
          const float dy = so[(0 * DG + grp) * PL + r];
          const float dx = so[(1 * DG + grp) * PL + r];
          const float m = so[(2 * DG + grp) * PL + r];
          const float py = (py_f[j] + (float)ti) + dy;
          const float px = (px_f[j] + (float)tj) + dx;
          const bool inside = (py > -1.f) && (py < (float)H) && (px > -1.f) && (px < (float)W);
          const float fy = floorf(py), fx = floorf(px);
          const int y0 = (int)fminf(fmaxf(fy, -2.f), 1.0e6f), x0 = (int)fminf(fmaxf(fx, -2.f), 1.0e6f);
          const float ly = py - fy, lx = px - fx;
          const bool vy0 = inside && y0 >= 0, vy1 = inside && y0 + 1 <= H - 1;
          const bool vx0 = x0 >= 0, vx1 = x0 + 1 <= W - 1;
          wgt[jj][0] = (vy0 && vx0) ? m * (1.f - ly) * (1.f - lx) : 0.f;
          wgt[jj][1] = (vy0 && vx1) ? m * (1.f - ly) * lx : 0.f;
          wgt[jj][2] = (vy1 && vx0) ? m * ly * (1.f - lx) : 0.f;
          wgt[jj][3] = (vy1 && vx1) ? m * ly * lx : 0.f;
          // Unconditional loads: an out-of-image corner reads a clamped in-image address with
          // weight 0 (no divergent branches around the four LDGs); 32-bit element offsets.
          const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
          const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1);
          const uint32_t b00 = (uint32_t)(cy0 * W + cx0) * CH + l * CPI;
          const uint32_t sxo = (uint32_t)(cx1 - cx0) * CH, syo = (uint32_t)((cy1 - cy0) * W) * CH;
          RawVec<XT, CPI>::ld(xn + b00, raw[jj][0]);
          RawVec<XT, CPI>::ld(xn + (uint32_t)(b00 + sxo), raw[jj][1]);
          RawVec<XT, CPI>::ld(xn + (uint32_t)(b00 + syo), raw[jj][2]);
          RawVec<XT, CPI>::ld(xn + (uint32_t)(b00 + syo + sxo), raw[jj][3]);
        }
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
          const int r = rbase + (j0 + jj) * Cfg::ROWS_PER_PASS;
          float v[CPI];
#pragma unroll
          for (int e = 0; e < CPI; ++e)
            v[e] = wgt[jj][0] * RawVec<XT, CPI>::get(raw[jj][0], e) + wgt[jj][1] * RawVec<XT, CPI>::get(raw[jj][1], e) +
                   wgt[jj][2] * RawVec<XT, CPI>::get(raw[jj][2], e) + wgt[jj][3] * RawVec<XT, CPI>::get(raw[jj][3], e);
          uint32_t hi[CPI / 2];
#pragma unroll
          for (int e = 0; e < CPI / 2; ++e) hi[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
          const uint32_t dst = aStage + sw128_offset(r, l * CPI * 2);
          if (CPI == 8) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst), "r"(hi[0]), "r"(hi[1]), "r"(hi[CPI / 2 - 2]), "r"(hi[CPI / 2 - 1]));
          else          asm volatile("st.shared.v2.b32 [%0], {%1, %2};\n" ::"r"(dst), "r"(hi[0]), "r"(hi[1]));
          if (SPLIT) {
            uint32_t lo[CPI / 2];
#pragma unroll
            for (int e = 0; e < CPI / 2; ++e)
              lo[e] = pack_bf16x2(v[2 * e] - bf16lo_to_f32(hi[e]), v[2 * e + 1] - bf16hi_to_f32(hi[e]));
            const uint32_t dl = dst + A_TILE_BYTES;
            if (CPI == 8) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dl), "r"(lo[0]), "r"(lo[1]), "r"(lo[CPI / 2 - 2]), "r"(lo[CPI / 2 - 1]));
            else          asm volatile("st.shared.v2.b32 [%0], {%1, %2};\n" ::"r"(dl), "r"(lo[0]), "r"(lo[1]));
          }
        }
      }
      fence_proxy_async_smem();  // generic-proxy writes (A tile, cp.async'd B) -> visible to the tensor core
      __syncthreads();
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        const uint32_t bStage = sB + (it % NB_STAGES) * SM::B_STAGE;
        const uint64_t a_hi = umma_desc_sw128_kmajor(aStage), b_hi = umma_desc_sw128_kmajor(bStage);
#pragma unroll
        for (int k = 0; k < CH / 16; ++k) {
          umma_bf16(tmem_d, a_hi + 2 * k, b_hi + 2 * k, IDESC, (tap | k) != 0);
          if (SPLIT) {
            const uint64_t a_lo = umma_desc_sw128_kmajor(aStage + A_TILE_BYTES);
            const uint64_t b_lo = umma_desc_sw128_kmajor(bStage + B_TILE_BYTES);
            umma_bf16(tmem_d, a_lo + 2 * k, b_hi + 2 * k, IDESC, 1);
            umma_bf16(tmem_d, a_hi + 2 * k, b_lo + 2 * k, IDESC, 1);
          }
        }
        umma_commit(bar0 + 8 * s);
      }
    }

    // ---- epilogue: TMEM -> registers -> (+bias) -> NHWC store -----------------------------
    {
      const int last = it - 1;
      mbar_wait(bar0 + 8 * (last & 1), (last >> 1) & 1);
      tc_fence_after();
      const int q = warp & 3, half = warp >> 2;
      uint32_t acc[32];
      tmem_ld_32x32(tmem_d + ((uint32_t)(q * 32) << 16) + half * 32, acc);
      tmem_ld_wait();
      const int r = q * 32 + (tid & 31);
      const int pix = pix0 + r;
      if (pix < HW) {
        XT* op = out + (size_t)n * os_n + (size_t)pix * CH + half * 32;
        float f[32];
#pragma unroll
        for (int e = 0; e < 32; ++e)
          f[e] = __uint_as_float(acc[e]) + (bias ? to_f32<XT>(bias[half * 32 + e]) : 0.f);
        if (sizeof(XT) == 2) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 u;
            u.x = pack_bf16x2(f[e], f[e + 1]); u.y = pack_bf16x2(f[e + 2], f[e + 3]);
            u.z = pack_bf16x2(f[e + 4], f[e + 5]); u.w = pack_bf16x2(f[e + 6], f[e + 7]);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(op) + e) = u;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(op) + e) = make_float4(f[e], f[e + 1], f[e + 2], f[e + 3]);
        }
      }
      tc_fence_before();  // order the TMEM reads before the next tile's first (overwriting) MMA
    }
  }

  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_d);
}

bool nhwc_dense(const int64_t s[4], int c, int h, int w) {
  return s[1] == 1 && s[3] == c && s[2] == (int64_t)w * c && s[0] >= (int64_t)h * w * c;
}

template <typename XT, bool SPLIT, int DG>
int launch_tc(const void* x, const int64_t* xs, const float* offset, const float* mask, const void* weight,
              const void* bias, void* out, const int64_t* os, int n, int h, int w, void* workspace,
              unsigned flags, cudaStream_t st) {
  using SM = TcSmem<SPLIT, DG>;
  int rc = 0;
  if (!(flags & EAVSR_DCN_WS_PACKED)) {
    dcn_pack_weight<XT, SPLIT><<<(TAPS * CH * CH + 255) / 256, 256, 0, st>>>((const XT*)weight, (uint8_t*)workspace);
    rc = check_launch("dcn_forward(pack)");
    if (rc) return rc;
  }
  const int HW = h * w;
  const int tiles_per_img = (HW + TILE_M - 1) / TILE_M;
  const int total = tiles_per_img * n;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int per_sm = SPLIT ? 1 : 2;
  const int grid = total < sms * per_sm ? total : sms * per_sm;
  const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(offset) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(mask) & 15u) == 0);
  if constexpr (!SPLIT) {
    if (!(flags & (EAVSR_DCN_FORCE_V1 | EAVSR_DCN_FORCE_WS)) && (long long)h * w <= (DG == 16 ? (1ll << 23) : (1ll << 24))) {
      // third generation: shared-memory window gather (bf16)
      const bool vecw = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(offset) & 15u) == 0) &&
                        ((reinterpret_cast<uintptr_t>(mask) & 15u) == 0);
      const int tiles_x = (w + win::TW - 1) / win::TW, tiles_y = (h + win::TH - 1) / win::TH;
      const int tpi = tiles_x * tiles_y, tot = tpi * n;
      const int g3 = tot < sms ? tot : sms;
      const bool b16 = !(flags & EAVSR_DCN_BLEND_FP32);
      if constexpr (DG == 8) {
        // fourth generation (dcn_fwd_win2.cuh): TMA-staged offsets, one pixel per lane.  Needs tensor maps over the
        // offset / mask tensors, i.e. 16-byte aligned rows (w % 4 == 0)
        if (vecw && !(flags & EAVSR_DCN_FORCE_WIN1)) {
          CUtensorMap tmo, tmm;
          const unsigned long long HWb = (unsigned long long)h * w * 4;
          const unsigned long long od[5] = {(unsigned long long)w, (unsigned long long)h, 18, 8, (unsigned long long)n};
          const unsigned long long os_[4] = {(unsigned long long)w * 4, HWb, 18 * HWb, 144 * HWb};
          const unsigned long long md[5] = {(unsigned long long)w, (unsigned long long)h, 9, 8, (unsigned long long)n};
          const unsigned long long ms_[4] = {(unsigned long long)w * 4, HWb, 9 * HWb, 72 * HWb};
          const unsigned box[5] = {win2::TW, win2::TH, 1, 8, 1};
          if (encode_tensor_map(&tmo, EAVSR_F32, 5, offset, od, os_, box, 0) &&
              encode_tensor_map(&tmm, EAVSR_F32, 5, mask, md, ms_, box, 0)) {
            // fifth generation (dcn_fwd_win3.cuh): A tile in tensor memory, window by one 4-D tensor load of x
            CUtensorMap tmx;
            const unsigned long long xd[4] = {(unsigned long long)CH, (unsigned long long)w, (unsigned long long)h,
                                              (unsigned long long)n};
            const unsigned long long xst[3] = {(unsigned long long)CH * 2, (unsigned long long)w * CH * 2,
                                               (unsigned long long)xs[0] * 2};
            const unsigned xbox[4] = {CH, win3::WW, win3::WH, 1};
            if (!(flags & EAVSR_DCN_FORCE_WIN2) && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 &&
                (reinterpret_cast<uintptr_t>(bias) & 15u) == 0 && tot < (1 << 22) &&
                encode_tensor_map(&tmx, EAVSR_BF16, 4, x, xd, xst, xbox, 0)) {
              auto k3 = b16 ? win3::dcn_fwd_win3_kernel<true> : win3::dcn_fwd_win3_kernel<false>;
              cudaError_t e3 = cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, win3::Smem::DYN);
              if (e3 != cudaSuccess) { set_error("dcn_forward(win3): smem attr: %s", cudaGetErrorString(e3)); return EAVSR_ERR_CUDA; }
              k3<<<g3, win3::THREADS, win3::Smem::DYN, st>>>((const __nv_bfloat16*)x, tmx, tmo, tmm,
                                                           (const uint8_t*)workspace, (const __nv_bfloat16*)bias,
                                                           (__nv_bfloat16*)out, h, w, xs[0], os[0], tiles_x, tpi, tot);
              return check_launch("dcn_forward(win3)");
            }
            auto k2 = b16 ? win2::dcn_fwd_win2_kernel<true> : win2::dcn_fwd_win2_kernel<false>;
            cudaError_t e2 = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, win2::Smem::DYN);
            if (e2 != cudaSuccess) { set_error("dcn_forward(win2): smem attr: %s", cudaGetErrorString(e2)); return EAVSR_ERR_CUDA; }
            k2<<<g3, win2::THREADS, win2::Smem::DYN, st>>>((const __nv_bfloat16*)x, tmo, tmm, (const uint8_t*)workspace,
                                                         (const __nv_bfloat16*)bias, (__nv_bfloat16*)out, h, w, xs[0],
                                                         os[0], tiles_x, tpi, tot);
            return check_launch("dcn_forward(win2)");
          }
        }
      }
      void (*k)(const __nv_bfloat16*, const float*, const float*, const uint8_t*, const __nv_bfloat16*,
                __nv_bfloat16*, int, int, long long, long long, int, int, int, const __nv_bfloat16*,
                const __nv_bfloat16*, int) =
          vecw ? (b16 ? win::dcn_fwd_win_kernel<DG, true, true> : win::dcn_fwd_win_kernel<DG, true, false>)
               : (b16 ? win::dcn_fwd_win_kernel<DG, false, true> : win::dcn_fwd_win_kernel<DG, false, false>);
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, win::Cfg<DG>::DYN);
      if (e != cudaSuccess) { set_error("dcn_forward(win): smem attr: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
      k<<<g3, win::THREADS, win::Cfg<DG>::DYN, st>>>((const __nv_bfloat16*)x, offset, mask, (const uint8_t*)workspace,
                                                 (const __nv_bfloat16*)bias, (__nv_bfloat16*)out, h, w, xs[0], os[0],
                                                 tiles_x, tpi, tot, nullptr, nullptr, CH);
      return check_launch("dcn_forward(win)");
    }
  }
  if constexpr (DG <= 8) {
    if (!(flags & EAVSR_DCN_FORCE_V1)) {       // warp-specialised second-generation kernel
      using WS = ws::Smem<SPLIT>;
      auto kv = ws::dcn_fwd_ws_kernel<XT, SPLIT, DG, true>;
      auto ks = ws::dcn_fwd_ws_kernel<XT, SPLIT, DG, false>;
      auto k = vec ? kv : ks;
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, WS::DYN_BYTES);
      if (e != cudaSuccess) { set_error("dcn_forward(ws): smem attr: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
      k<<<grid, ws::THREADS, WS::DYN_BYTES, st>>>((const XT*)x, offset, mask, (const uint8_t*)workspace,
                                                   (const XT*)bias, (XT*)out, h, w, xs[0], os[0], tiles_per_img, total);
      return check_launch("dcn_forward(ws)");
    }
  }
  auto kv = dcn_fwd_tc_kernel<XT, SPLIT, DG, true>;
  auto ks = dcn_fwd_tc_kernel<XT, SPLIT, DG, false>;
  auto k = vec ? kv : ks;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::DYN_BYTES);
  if (e != cudaSuccess) { set_error("dcn_forward(tc): smem attr: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
  k<<<grid, TC_THREADS, SM::DYN_BYTES, st>>>((const XT*)x, offset, mask, (const uint8_t*)workspace, (const XT*)bias,
                                            (XT*)out, h, w, xs[0], os[0], tiles_per_img, total);
  return check_launch("dcn_forward(tc)");
}

template <typename XT, bool SPLIT>
int launch_tc_dg(int dg, const void* x, const int64_t* xs, const float* offset, const float* mask,
                 const void* weight, const void* bias, void* out, const int64_t* os, int n, int h, int w,
                 void* workspace, unsigned flags, cudaStream_t st) {
  switch (dg) {
#ifndef EAVSR_ONLY_DG8   // (development builds: tools/abl_build.sh compiles the deform_groups = 8 kernels only)
    case 1: return launch_tc<XT, SPLIT, 1>(x, xs, offset, mask, weight, bias, out, os, n, h, w, workspace, flags, st);
    case 2: return launch_tc<XT, SPLIT, 2>(x, xs, offset, mask, weight, bias, out, os, n, h, w, workspace, flags, st);
    case 4: return launch_tc<XT, SPLIT, 4>(x, xs, offset, mask, weight, bias, out, os, n, h, w, workspace, flags, st);
    case 16: return launch_tc<XT, SPLIT, 16>(x, xs, offset, mask, weight, bias, out, os, n, h, w, workspace, flags, st);
#endif
    case 8: return launch_tc<XT, SPLIT, 8>(x, xs, offset, mask, weight, bias, out, os, n, h, w, workspace, flags, st);
  }
  set_error("dcn_forward(tc): deform_groups %d", dg);
  return EAVSR_ERR_INVALID;
}

int fill_geom(DcnGeom& g, int n, int cin, int h, int w, int cout, int kh, int kw, int sh, int sw, int ph, int pw,
              int dh, int dw, int groups, int dg, const char* who) {
  EAVSR_REQUIRE(n > 0 && cin > 0 && h > 0 && w > 0 && cout > 0, "%s: empty tensor", who);
  EAVSR_REQUIRE(kh > 0 && kw > 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0 && ph >= 0 && pw >= 0, "%s: bad conv geometry", who);
  EAVSR_REQUIRE(groups > 0 && dg > 0 && cin % groups == 0 && cout % groups == 0 && cin % dg == 0,
                "%s: channels (%d,%d) not divisible by groups %d / deform_groups %d", who, cin, cout, groups, dg);
  EAVSR_REQUIRE(cout / groups <= 256, "%s: cout/groups > 256 not supported", who);
  g = DcnGeom{n, cin, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw, groups, dg, 0, 0};
  g.HO = (h + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1;
  g.WO = (w + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
  EAVSR_REQUIRE(g.HO > 0 && g.WO > 0, "%s: empty output (%d x %d)", who, g.HO, g.WO);
  return EAVSR_OK;
}

}  // namespace
}  // namespace eavsr

using namespace eavsr;

extern "C" size_t eavsr_dcn_forward_workspace(int cin, int cout, int kh, int kw, int groups, int deform_groups,
                                              int dtype) {
  (void)groups; (void)deform_groups;
  if (cin != CH || cout != CH || kh != 3 || kw != 3) return 0;
  return (size_t)TAPS * B_TILE_BYTES * (dtype == EAVSR_F32 ? 2 : 1);
}

extern "C" int eavsr_dcn_forward_uses_tensor_cores(const int64_t x_strides[4], const int64_t out_strides[4],
                                                   int cin, int cout, int kh, int kw, int sh, int sw, int ph,
                                                   int pw, int dh, int dw, int groups, int deform_groups,
                                                   unsigned flags) {
  if (flags & EAVSR_DCN_FORCE_GENERIC) return 0;
  const bool cfg = cin == CH && cout == CH && kh == 3 && kw == 3 && sh == 1 && sw == 1 && ph == 1 && pw == 1 &&
                   dh == 1 && dw == 1 && groups == 1 &&
                   (deform_groups == 1 || deform_groups == 2 || deform_groups == 4 || deform_groups == 8 ||
                    deform_groups == 16);
  if (!cfg) return 0;
  // NHWC dense in h,w,c (batch stride free)
  const bool lx = x_strides[1] == 1 && x_strides[3] == CH && x_strides[2] % CH == 0;
  const bool lo = out_strides[1] == 1 && out_strides[3] == CH && out_strides[2] % CH == 0;
  return lx && lo;
}

extern "C" int eavsr_dcn_forward(const void* x, const int64_t x_strides[4], const float* offset, const float* mask,
                                 const void* weight, const void* bias, void* out, const int64_t out_strides[4],
                                 int n, int cin, int h, int w, int cout, int kh, int kw, int sh, int sw, int ph,
                                 int pw, int dh, int dw, int groups, int deform_groups, int dtype, void* workspace,
                                 size_t workspace_bytes, unsigned flags, void* stream) {
  EAVSR_REQUIRE(x && offset && mask && weight && out && x_strides && out_strides, "dcn_forward: null pointer");
  EAVSR_REQUIRE(dtype == EAVSR_F32 || dtype == EAVSR_BF16, "dcn_forward: bad dtype %d", dtype);
  DcnGeom g;
  int rc = fill_geom(g, n, cin, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw, groups, deform_groups, "dcn_forward");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esz = dtype == EAVSR_F32 ? 4 : 2;
  bool tc = eavsr_dcn_forward_uses_tensor_cores(x_strides, out_strides, cin, cout, kh, kw, sh, sw, ph, pw, dh, dw,
                                                groups, deform_groups, flags) != 0;
  if (tc) {
    // the gather assumes rows of W*64 elements and 16-byte aligned pixels
    tc = nhwc_dense(x_strides, cin, h, w) && nhwc_dense(out_strides, cout, h, w) &&
         ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) &&
         ((x_strides[0] * esz) % 16 == 0) && ((out_strides[0] * esz) % 16 == 0) &&
         (!bias || (reinterpret_cast<uintptr_t>(bias) & 3u) == 0) && ((long long)h * w * CH < (1ll << 31));
  }
  if (tc) {
    const size_t need = eavsr_dcn_forward_workspace(cin, cout, kh, kw, groups, deform_groups, dtype);
    EAVSR_REQUIRE(workspace && workspace_bytes >= need && ((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0),
                  "dcn_forward: tensor-core path needs a 16-byte aligned workspace of %zu bytes (got %zu)", need,
                  workspace_bytes);
    if (dtype == EAVSR_F32)
      return launch_tc_dg<float, true>(deform_groups, x, x_strides, offset, mask, weight, bias, out, out_strides, n, h,
                                       w, workspace, flags, st);
    return launch_tc_dg<__nv_bfloat16, false>(deform_groups, x, x_strides, offset, mask, weight, bias, out,
                                              out_strides, n, h, w, workspace, flags, st);
  }
  if (dtype == EAVSR_F32)
    return dcn_forward_generic<float>(x, x_strides, offset, mask, weight, bias, out, out_strides, g, st);
  return dcn_forward_generic<__nv_bfloat16>(x, x_strides, offset, mask, weight, bias, out, out_strides, g, st);
}

namespace eavsr {
namespace {
template <int DG>
int launch_affine(const void* x, const int64_t* xs, const void* affine, const void* affine_bias, const void* weight,
                  const void* bias, void* out, const int64_t* os, int n, int h, int w, void* workspace, unsigned flags,
                  cudaStream_t st) {
  int rc = 0;
  if (!(flags & EAVSR_DCN_WS_PACKED)) {
    dcn_pack_weight<__nv_bfloat16, false><<<(TAPS * CH * CH + 255) / 256, 256, 0, st>>>((const __nv_bfloat16*)weight,
                                                                                      (uint8_t*)workspace);
    rc = check_launch("dcn_affine_forward(pack)");
    if (rc) return rc;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles_x = (w + win::TW - 1) / win::TW, tiles_y = (h + win::TH - 1) / win::TH;
  const int tpi = tiles_x * tiles_y, tot = tpi * n;
  const int grid = tot < sms ? tot : sms;
  const bool b16 = !(flags & EAVSR_DCN_BLEND_FP32);
  auto k = b16 ? win::dcn_fwd_win_kernel<DG, true, true, true> : win::dcn_fwd_win_kernel<DG, true, false, true>;
  using C = win::Cfg<DG, true>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::DYN);
  if (e != cudaSuccess) { set_error("dcn_affine_forward: smem attr: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
  k<<<grid, win::THREADS, C::DYN, st>>>((const __nv_bfloat16*)x, nullptr, nullptr, (const uint8_t*)workspace,
                                        (const __nv_bfloat16*)bias, (__nv_bfloat16*)out, h, w, xs[0], os[0], tiles_x,
                                        tpi, tot, (const __nv_bfloat16*)affine, (const __nv_bfloat16*)affine_bias,
                                        (int)os[3]);
  return check_launch("dcn_affine_forward");
}
}  // namespace
}  // namespace eavsr

extern "C" int eavsr_dcn_affine_forward(const void* x, const int64_t x_strides[4], const void* affine,
                                        const void* affine_bias, const void* weight, const void* bias, void* out,
                                        const int64_t out_strides[4], int n, int h, int w, int deform_groups,
                                        int dtype, void* workspace, size_t workspace_bytes, unsigned flags,
                                        void* stream) {
  EAVSR_REQUIRE(x && affine && weight && out && x_strides && out_strides, "dcn_affine_forward: null pointer");
  EAVSR_REQUIRE(n > 0 && h > 0 && w > 0, "dcn_affine_forward: empty tensor");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool ok = dtype == EAVSR_BF16 &&
                  deform_groups == 8 &&   // 15*dg bf16 per pixel must be a multiple of 16 bytes
                  nhwc_dense(x_strides, CH, h, w) && al16(x) && al16(out) &&
                  // out may be a 64-channel slice of a wider NHWC buffer (pixel stride >= 64, multiple of 8)
                  out_strides[1] == 1 && out_strides[3] >= CH && out_strides[3] % 8 == 0 && out_strides[3] < (1 << 20) &&
                  out_strides[2] == (int64_t)w * out_strides[3] && out_strides[0] >= (int64_t)h * w * out_strides[3] &&
                  al16(affine) && (x_strides[0] * 2) % 16 == 0 && (out_strides[0] * 2) % 16 == 0 &&
                  (long long)h * w <= (1ll << 24) && (!bias || (reinterpret_cast<uintptr_t>(bias) & 3u) == 0);
  if (!ok) {
    set_error("dcn_affine_forward: only bf16 dense NHWC 64->64 with deform_groups = 8 is fused");
    return EAVSR_ERR_UNSUPPORTED;
  }
  const size_t need = eavsr_dcn_forward_workspace(CH, CH, 3, 3, 1, deform_groups, dtype);
  EAVSR_REQUIRE(workspace && workspace_bytes >= need && al16(workspace),
                "dcn_affine_forward: needs a 16-byte aligned workspace of %zu bytes (got %zu)", need, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  return launch_affine<8>(x, x_strides, affine, affine_bias, weight, bias, out, out_strides, n, h, w, workspace, flags, st);
}

namespace eavsr {
size_t dcn_backward_tc_workspace();
bool dcn_backward_tc_eligible(const int64_t* gs, const int64_t* xs, const int64_t* gxs, const DcnGeom& g, bool has_gx);
int dcn_backward_tc(const void* gout, const int64_t* gs, const void* x, const int64_t* xs, const float* offset,
                    const float* mask, const void* weight, float* gx32, const int64_t* gxs, float* goffset,
                    float* gmask, float* gweight32, float* gbias32, const DcnGeom& g, void* workspace, unsigned which,
                    cudaStream_t st);
}  // namespace eavsr

extern "C" size_t eavsr_dcn_backward_workspace(int cin, int cout, int kh, int kw, int groups, int deform_groups,
                                               int dtype) {
  if (cin != CH || cout != CH || kh != 3 || kw != 3 || groups != 1 || dtype != EAVSR_BF16) return 0;
  if (!(deform_groups == 1 || deform_groups == 2 || deform_groups == 4 || deform_groups == 8 || deform_groups == 16))
    return 0;
  return dcn_backward_tc_workspace();
}

extern "C" int eavsr_dcn_backward(const void* gout, const int64_t gout_strides[4], const void* x,
                                  const int64_t x_strides[4], const float* offset, const float* mask,
                                  const void* weight, float* gx32, const int64_t gx_strides[4], float* goffset,
                                  float* gmask, float* gweight32, float* gbias32, int n, int cin, int h, int w,
                                  int cout, int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int groups,
                                  int deform_groups, int dtype, void* workspace, size_t workspace_bytes,
                                  unsigned flags, void* stream) {
  EAVSR_REQUIRE(gout && x && offset && mask && weight && gout_strides && x_strides, "dcn_backward: null pointer");
  EAVSR_REQUIRE(!gx32 || gx_strides, "dcn_backward: gx32 without strides");
  EAVSR_REQUIRE(dtype == EAVSR_F32 || dtype == EAVSR_BF16, "dcn_backward: bad dtype %d", dtype);
  DcnGeom g;
  int rc = fill_geom(g, n, cin, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw, groups, deform_groups, "dcn_backward");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == EAVSR_F32)
    return dcn_backward_generic<float>(gout, gout_strides, x, x_strides, offset, mask, weight, gx32, gx_strides,
                                       goffset, gmask, gweight32, gbias32, g, st);
  unsigned which = 0;   // bit 0: data gradients, bit 1: weight gradient on the tcgen05 kernels
  const size_t need = eavsr_dcn_backward_workspace(cin, cout, kh, kw, groups, deform_groups, dtype);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!(flags & EAVSR_DCN_FORCE_GENERIC) && need && workspace && workspace_bytes >= need && al16(workspace) &&
      al16(gout) && al16(x) && (!gx32 || al16(gx32)) &&
      dcn_backward_tc_eligible(gout_strides, x_strides, gx_strides, g, gx32 != nullptr)) {
    if (!(flags & EAVSR_DCN_BWD_GENERIC_DATA)) which |= 1u;
    if (!(flags & EAVSR_DCN_BWD_GENERIC_WEIGHT)) which |= 2u;
  }
  if (which) {
    rc = dcn_backward_tc(gout, gout_strides, x, x_strides, offset, mask, weight, gx32, gx_strides, goffset, gmask,
                         gweight32, gbias32, g, workspace, which, st);
    if (rc) return rc;
    gbias32 = nullptr;
  }
  const bool tc_data = (which & 1u) != 0, tc_w = (which & 2u) != 0;
  return dcn_backward_generic<__nv_bfloat16>(gout, gout_strides, x, x_strides, offset, mask, weight,
                                             tc_data ? nullptr : gx32, gx_strides, tc_data ? nullptr : goffset,
                                             tc_data ? nullptr : gmask, tc_w ? nullptr : gweight32, gbias32, g, st);
}
