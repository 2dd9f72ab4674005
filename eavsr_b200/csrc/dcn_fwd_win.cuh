// Third-generation tcgen05 DCNv2 forward (bf16 features, deform_groups 1..16): the gather reads a
// shared-memory WINDOW of x instead of going through L1.
//
// Why (ncu, profiles/r1_dcn_fwd_ws_ncu.txt): with per-group offsets every lane of a gather request
// touches its own 128-byte line but uses 16 bytes of it -- 23 sectors per request, L1 throughput
// 55-70 %, L1 hit rate 14-38 %.  The 4 608 gathered bytes per output pixel are 36x the 128 bytes of x
// that pixel owns, i.e. the reuse is there, L1 just cannot serve it at 16-byte granularity.
// Shared memory can: with pixel-major rows of 128 B the 8 lanes of a quarter-warp are the 8 channel
// chunks, i.e. 8 distinct 16-byte bank groups whatever pixels they hit -> conflict-free LDS.128.
//
//   * tile = 8 x 16 output pixels (M = 128); window = the tile plus a 5-pixel apron (18 x 26 pixels,
//     58.5 KB; 4 pixels for deform_groups = 16 and the fused-offset variant), double buffered, filled by
//     one elected lane with cp.async.bulk row copies (TMA engine) for the NEXT tile while this one is
//     gathered.  Out-of-image cells are zeroed once per border tile, so an outside corner reads 0 -- DCNv2's
//     zero padding and its "p <= -1 or p >= size -> 0" rule at once, without per-corner tests.
//   * a sample whose four corners are not all inside the window (|offset| > ~4 px) falls back to a
//     global-memory gather with explicit validity -- correct for any offset, fast for real ones.
//   * 16 producer warps (2 items per thread and tap, advanced in lock step) + 1 MMA / loader warp;
//     bf16x2 HFMA2 blend (template BLEND16) or fp32 blend; per-warp cp.async staging of the offset / mask
//     planes (3-deep ring), or -- AFF -- of the raw affine block once per tile; A ring of 3 stages; per-tap
//     weight tiles through a 2-stage bulk-copy ring refilled behind the queued MMAs; 2 TMEM accumulators.
#pragma once
#include "common.cuh"

#ifndef EAVSR_ABL
#define EAVSR_ABL 0   // development only (tools/abl_build.sh): bit mask of pipeline pieces to leave out for TIMING
#endif               // 1 fence, 2 offset prefetch, 4 blend, 8 window loads, 16 A-tile stores, 32 coordinate math

namespace eavsr {
namespace win {

constexpr int PWARPS = 16;
constexpr int THREADS = (PWARPS + 1) * 32;   // 544
constexpr int TH = 8, TW = 16;               // tile
constexpr int CH = 64, TAPS = 9;
constexpr int A_TILE = 128 * CH * 2;         // 16 KB
constexpr int B_TILE = CH * CH * 2;          // 8 KB
#ifndef EAVSR_WIN_NSA
#define EAVSR_WIN_NSA 3
#endif
#ifndef EAVSR_WIN_NSB
#define EAVSR_WIN_NSB 2
#endif
#ifndef EAVSR_WIN_NOB
#define EAVSR_WIN_NOB 3
#endif
constexpr int NSA = EAVSR_WIN_NSA, NSB = EAVSR_WIN_NSB;   // A-stage ring, weight-tile ring
constexpr int PLW = 8;                       // staged offset plane: 8 pixels, XOR-swizzled (no padding)
constexpr int TMEM_COLS = 128;

// Per-deform_groups configuration.  deform_groups <= 8: one 16-byte chunk (8 channels) belongs to one
// group, 5-pixel apron, 3 offset buffers per warp.  deform_groups = 16 (BASELINE config 2): a chunk holds
// two groups of 4 channels (two samples with their own offsets per 16 bytes of the A tile, 8-byte corner
// loads) and 48 offset/mask planes per tap; the planes take twice the room, so the apron shrinks to 4
// pixels and the offset ring to 2 buffers to stay inside 227 KB.
// Offsets/masks come from HBM with ~1.5 us latency and each (tile, tap) needs 12 KB of them; with a
// prefetch distance of 2 taps the first versions had only ~24 KB per SM in flight (Little: 2.4 TB/s
// for the whole chip = the ~100 us floor every earlier kernel hit).  3 buffers -> distance 2 (deeper
// did not pay: the A ring depth did).
// AFF (fused offset generation): the kernel reads the (n, h, w, 15*DG) bf16 output of
// the offset-generating convolution (4*DG transform, 2*DG translation, 9*DG mask logits per pixel) instead
// of fp32 offset / mask planes: a warp stages the 8 x 15*DG*2 bytes of its pixels once per TILE (double
// buffered), keeps the affine part in registers and expands offset = T*R - R + t and sigmoid(logit) per tap.
template <int DG, bool AFF = false> struct Cfg {
  static constexpr int NSUB = DG == 16 ? 2 : 1;           // samples (groups) per 16-byte chunk
  static constexpr int PAD = (DG == 16 || AFF) ? 4 : 5;
  static constexpr int WH = TH + 2 * PAD, WW = TW + 2 * PAD;   // 18 x 26 (16 x 24) window
  static constexpr int WIN_BYTES = WH * WW * 128;              // 59 904 (49 152)
  static constexpr int NOB = (DG == 16 || AFF) ? 2 : EAVSR_WIN_NOB;
  static constexpr int MAX_PLANES = DG == 16 ? 48 : 24;
  static constexpr int APX = 15 * DG * 2;                      // AFF: bytes per pixel of the affine block
  static constexpr int OFF_WARP_BUF = AFF ? PLW * APX : MAX_PLANES * PLW * 4;    // 768 B (1 536 B; AFF: 1 920 B)
  static constexpr int WIN_OFF = 0;
  static constexpr int A_OFF = WIN_OFF + 2 * WIN_BYTES;        // multiple of 1024
  static constexpr int B_OFF = A_OFF + NSA * A_TILE;
  static constexpr int OFFS_OFF = B_OFF + NSB * B_TILE;
  static constexpr int BAR_OFF = OFFS_OFF + PWARPS * NOB * OFF_WARP_BUF;
  // full[NSA], empty[NSA], bfull[NSB], accf[2], acce[2], winf[2], wine[2]; tmem slot
  static constexpr int NBARS = 2 * NSA + NSB + 8;
  static constexpr int ABIAS_OFF = BAR_OFF + NBARS * 8 + 16;   // AFF: 15*DG fp32 biases
  static constexpr int TOTAL = ABIAS_OFF + (AFF ? 15 * DG * 4 : 0);
  static constexpr int DYN = TOTAL + 1024;
  static_assert(A_OFF % 1024 == 0, "A stages must be 1024-byte aligned for the 128B swizzle");
  static_assert(DYN <= 232448, "shared memory budget");
  static_assert(!AFF || (DG <= 8 && APX % 16 == 0), "fused offsets: deform_groups <= 8");
};
using Smem = Cfg<8>;                         // layout shared by deform_groups 1, 2, 4, 8

__device__ __forceinline__ uint32_t hmul2_bf16(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;\n" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t hfma2_bf16(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;\n" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

template <int DG, bool VEC_OFF, bool BLEND16, bool AFF = false>
__global__ void __launch_bounds__(THREADS, 1)
dcn_fwd_win_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ offset,
                   const float* __restrict__ mask, const uint8_t* __restrict__ wpacked,
                   const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W,
                   long long xs_n, long long os_n, int tiles_x, int tiles_per_img, int total_tiles,
                   const __nv_bfloat16* __restrict__ aff = nullptr,
                   const __nv_bfloat16* __restrict__ aff_bias = nullptr, int os_pix = CH) {
  using C = Cfg<DG, AFF>;
  using Smem = Cfg<DG, AFF>;
  constexpr int NSUB = C::NSUB, PAD = C::PAD, WH = C::WH, WW = C::WW, WIN_BYTES = C::WIN_BYTES;
  constexpr int NOB = C::NOB, OFF_WARP_BUF = C::OFF_WARP_BUF;
  constexpr int NPLANES = 3 * DG;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t sWin = sbase + Smem::WIN_OFF, sA = sbase + Smem::A_OFF, sB = sbase + Smem::B_OFF;
  const uint32_t bars = sbase + Smem::BAR_OFF;
  const uint32_t bar_full = bars, bar_empty = bar_full + NSA * 8, bar_bfull = bar_empty + NSA * 8;
  const uint32_t bar_accf = bar_bfull + NSB * 8, bar_acce = bar_accf + 16;
  const uint32_t bar_winf = bar_acce + 16, bar_wine = bar_winf + 16;
  const uint32_t tmem_slot_addr = bar_wine + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Smem::BAR_OFF + Smem::NBARS * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = H * W;
  if (tid == 0) {
    for (int s = 0; s < NSA; ++s) { mbar_init(bar_full + 8 * s, PWARPS); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < NSB; ++s) mbar_init(bar_bfull + 8 * s, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accf + 8 * b, 1);
      mbar_init(bar_acce + 8 * b, PWARPS);
      mbar_init(bar_winf + 8 * b, 1);
      mbar_init(bar_wine + 8 * b, PWARPS);
    }
    fence_mbar_init();
  }
  if constexpr (AFF) {
    float* sb = reinterpret_cast<float*>(smem + Smem::ABIAS_OFF);
    for (int i = tid; i < 15 * DG; i += THREADS) sb[i] = aff_bias ? __bfloat162float(aff_bias[i]) : 0.f;
  }
  // window cells that lie outside the image are never loaded; zero both buffers once so that whatever
  // they hold later (stale x) is finite and their 0 weights really give 0
  for (int i = tid; i < 2 * WIN_BYTES / 16; i += THREADS)
    *reinterpret_cast<uint4*>(smem + Smem::WIN_OFF + i * 16) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == PWARPS) tmem_alloc<TMEM_COLS>(tmem_slot_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  const int first = blockIdx.x;
  const int my_tiles = (total_tiles - first + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_iters = my_tiles * TAPS;
  constexpr uint32_t IDESC = umma_idesc_bf16(128, CH);

  auto tile_coords = [&](int tl, int& n, int& ty0, int& tx0) {
    const int tile = first + tl * (int)gridDim.x;
    n = tile / tiles_per_img;
    const int rem = tile - n * tiles_per_img;
    ty0 = (rem / tiles_x) * TH;
    tx0 = (rem % tiles_x) * TW;
  };

  if (warp == PWARPS) {
    // ============ MMA issuer + weight-tile loader + window loader (one lane) ============
    if (elect_one()) {
      auto issue_b = [&](int j) {
        const uint32_t bar = bar_bfull + 8 * (j % NSB);
        mbar_arrive_expect_tx(bar, B_TILE);
        bulk_g2s(sB + (j % NSB) * B_TILE, wpacked + (size_t)(j % TAPS) * B_TILE, B_TILE, bar);
      };
      auto load_window = [&](int tl) {
        int n, ty0, tx0;
        tile_coords(tl, n, ty0, tx0);
        const int wy0 = ty0 - PAD, wx0 = tx0 - PAD;
        const int gx0 = max(wx0, 0), gx1 = min(wx0 + WW, W);
        const int gy0 = max(wy0, 0), gy1 = min(wy0 + WH, H);
        const uint32_t bar = bar_winf + 8 * (tl & 1);
        const uint32_t row_bytes = (uint32_t)(gx1 - gx0) * 128u;
        mbar_arrive_expect_tx(bar, row_bytes * (uint32_t)(gy1 - gy0));
        const __nv_bfloat16* xn = x + (size_t)n * xs_n;
        const uint32_t dst0 = sWin + (tl & 1) * WIN_BYTES;
        for (int gy = gy0; gy < gy1; ++gy)
          bulk_g2s(dst0 + ((gy - wy0) * WW + (gx0 - wx0)) * 128, xn + ((size_t)gy * W + gx0) * CH, row_bytes, bar);
      };
      load_window(0);
      for (int j = 0; j < NSB && j < n_iters; ++j) issue_b(j);
      const uint64_t a_base = umma_desc_sw128_kmajor(sA), b_base = umma_desc_sw128_kmajor(sB);
      int tap = 0, tl = 0;
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % NSA, sb = it % NSB, buf = tl & 1;
        if (tap == 0 && tl >= 2) mbar_wait(bar_acce + 8 * buf, ((tl >> 1) - 1) & 1);
        mbar_wait(bar_full + 8 * s, (it / NSA) & 1);
        if (tap == 0 && tl + 1 < my_tiles) {               // producers have left tile tl-1: refill its window
          if (tl >= 1) mbar_wait(bar_wine + 8 * ((tl + 1) & 1), (((tl + 1) >> 1) - 1) & 1);
          load_window(tl + 1);
        }
        mbar_wait(bar_bfull + 8 * sb, (it / NSB) & 1);
        tc_fence_after();
        // descriptors differ from the stage-0 ones only in the 16-byte-granular start address field
        const uint64_t a_d = a_base + (uint64_t)((s * A_TILE) >> 4), b_d = b_base + (uint64_t)((sb * B_TILE) >> 4);
        const uint32_t d = tmem_d + buf * CH;
#pragma unroll
        for (int k = 0; k < CH / 16; ++k) umma_bf16(d, a_d + 2 * k, b_d + 2 * k, IDESC, (tap | k) != 0);
        umma_commit(bar_empty + 8 * s);
        if (tap == TAPS - 1) umma_commit(bar_accf + 8 * buf);
        // refill the weight ring behind the MMAs: the stage of it-1 is free once MMA(it-1) retired, and
        // MMA(it) is already queued, so the tensor pipe never waits for this thread
        if (it >= 1 && it - 1 + NSB < n_iters) {
          mbar_wait(bar_empty + 8 * ((it - 1) % NSA), ((it - 1) / NSA) & 1);
          issue_b(it - 1 + NSB);
        }
        if (++tap == TAPS) { tap = 0; ++tl; }
      }
    }
    __syncwarp();
  } else {
    // ============ producers: window gather -> blend -> swizzled A stage ============
    const int q = lane >> 3, l = lane & 7;
    const int wrow = warp >> 1, wcol = (warp & 1) * 8;       // this warp: tile row wrow, columns wcol..wcol+7
    const uint32_t offBase = sbase + Smem::OFFS_OFF + warp * NOB * OFF_WARP_BUF;
    const float* offF = reinterpret_cast<const float*>(smem + Smem::OFFS_OFF + warp * NOB * OFF_WARP_BUF);
    // column swizzle of the staged planes: groups 4..7 (12..15) swap the two 4-pixel halves, which makes
    // the 32 lanes (8 groups x 4 pixels) of one load hit 32 distinct banks without padding

    // Offset/mask prefetch stream, 3 taps ahead of the gather.  Everything that needs an integer
    // division (tile -> image / row / column) is done once per tile, not once per tap.
    struct TileRef { const float* ob; const float* mb; uint32_t ok; };
    // per-lane constants of the staging pattern: which plane / columns copy k of this lane moves
    constexpr int PER_PLANE = VEC_OFF ? 2 : 8;
    constexpr int NCOPY = (NPLANES * PER_PLANE + 31) / 32;
    uint32_t cp_rel[NCOPY], cp_step[NCOPY], cp_dst[NCOPY];
    int cp_col[NCOPY];
    bool cp_mask[NCOPY];
#pragma unroll
    for (int k = 0; k < NCOPY; ++k) {
      const int i = lane + 32 * k;
      const int plane = min(i / PER_PLANE, NPLANES - 1), e = i % PER_PLANE;
      const int comp = plane / DG, g = plane - comp * DG;
      cp_col[k] = (i < NPLANES * PER_PLANE) ? (VEC_OFF ? e * 4 : e) : (1 << 28);   // unused copy: never in range
      cp_mask[k] = comp == 2;
      cp_rel[k] = (comp < 2 ? (uint32_t)(g * TAPS * 2 + comp) : (uint32_t)(g * TAPS)) * (uint32_t)HW + (uint32_t)(cp_col[k] & 15);
      cp_step[k] = (comp < 2 ? 2u : 1u) * (uint32_t)HW;
      cp_dst[k] = (uint32_t)(plane * PLW + ((cp_col[k] & 15) ^ (((g >> 2) & 1) << 2))) * 4u;
    }
    auto make_ref = [&](int tl) {
      TileRef r;
      int n, ty0, tx0;
      tile_coords(tl, n, ty0, tx0);
      const int gy = ty0 + wrow, gx = tx0 + wcol;
      const size_t pix = (size_t)gy * W + gx;
      r.ob = offset + (size_t)n * DG * TAPS * 2 * HW + pix;
      r.mb = mask + (size_t)n * DG * TAPS * HW + pix;
      r.ok = 0;
#pragma unroll
      for (int k = 0; k < NCOPY; ++k) r.ok |= (gy < H && gx + cp_col[k] < W) ? (1u << k) : 0u;
      return r;
    };
    TileRef pref = make_ref(0);          // tile of the prefetch stream
    int p_tl = 0, p_tap = 0, p_ring = 0;
    auto prefetch_next = [&]() {         // stages (p_tl, p_tap) into buffer p_ring, then advances
      if (p_tl < my_tiles) {
        const uint32_t dst0 = offBase + p_ring * OFF_WARP_BUF;
#pragma unroll
        for (int k = 0; k < NCOPY; ++k) {
          if (!(EAVSR_ABL & 2) && (pref.ok & (1u << k))) {
            const float* src = (cp_mask[k] ? pref.mb : pref.ob) + (cp_rel[k] + (uint32_t)p_tap * cp_step[k]);
            if (VEC_OFF) cp_async_16(dst0 + cp_dst[k], src); else cp_async_4(dst0 + cp_dst[k], src);
          }
        }
        if (++p_ring == NOB) p_ring = 0;
        if (++p_tap == TAPS) {
          p_tap = 0;
          if (++p_tl < my_tiles) pref = make_ref(p_tl);
        }
      }
      cp_async_commit();
    };

    // AFF: the affine block of this warp's 8 pixels, once per tile, double buffered
    auto prefetch_tile = [&](int tl) {
      if (AFF && tl < my_tiles) {
        int n, ty0, tx0;
        tile_coords(tl, n, ty0, tx0);
        const int gy = ty0 + wrow, gx = tx0 + wcol;
        const uint32_t dst0 = offBase + (tl & 1) * OFF_WARP_BUF;
        if (gy < H) {
          const __nv_bfloat16* src = aff + (((size_t)n * H + gy) * W + gx) * (15 * DG);
          constexpr int CPX = C::APX / 16;                     // 16-byte chunks per pixel
          for (int i = lane; i < PLW * CPX; i += 32) {
            const int px = i / CPX, c = i - px * CPX;
            if (gx + px < W) cp_async_16(dst0 + px * C::APX + c * 16, src + px * (15 * DG) + c * 8);
          }
        }
      }
      cp_async_commit();
    };

    auto epilogue = [&](int tl) {
      const int buf = tl & 1;
      mbar_wait(bar_accf + 8 * buf, (tl >> 1) & 1);
      tc_fence_after();
      const int qd = warp & 3, cq = warp >> 2;                // TMEM lane quadrant, 16-column quarter
      uint32_t acc[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
          : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]),
            "=r"(acc[7]), "=r"(acc[8]), "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]),
            "=r"(acc[14]), "=r"(acc[15])
          : "r"(tmem_d + ((uint32_t)(qd * 32) << 16) + buf * CH + cq * 16)
          : "memory");
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce + 8 * buf);
      int n, ty0, tx0;
      tile_coords(tl, n, ty0, tx0);
      const int m = qd * 32 + lane;
      const int gy = ty0 + (m >> 4), gx = tx0 + (m & 15);
      if (gy < H && gx < W) {
        __nv_bfloat16* op = out + (size_t)n * os_n + ((size_t)gy * W + gx) * os_pix + cq * 16;
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(acc[e]) + (bias ? __bfloat162float(bias[cq * 16 + e]) : 0.f);
#pragma unroll
        for (int e = 0; e < 16; e += 8) {
          uint4 u;
          u.x = pack_bf16x2(f[e], f[e + 1]); u.y = pack_bf16x2(f[e + 2], f[e + 3]);
          u.z = pack_bf16x2(f[e + 4], f[e + 5]); u.w = pack_bf16x2(f[e + 6], f[e + 7]);
          *reinterpret_cast<uint4*>(op + e) = u;
        }
      }
    };

    if constexpr (AFF) {
      prefetch_tile(0);
      prefetch_tile(1);
    } else {
#pragma unroll
      for (int i = 0; i < NOB - 1; ++i) prefetch_next();
    }
    int it = 0, ring = 0, oring = 0;      // ring = it % NSA, ring_ph = parity of it / NSA, oring = it % NOB
    uint32_t ring_ph = 0;
    for (int tl = 0; tl < my_tiles; ++tl) {
      int n, ty0, tx0;
      tile_coords(tl, n, ty0, tx0);
      const __nv_bfloat16* xn = x + (size_t)n * xs_n;
      const int wy0 = ty0 - PAD, wx0 = tx0 - PAD;
      const uint32_t win = sWin + (tl & 1) * WIN_BYTES + l * 16;
      const int gy = ty0 + wrow;
      float pyb[2], pxb[2];                                   // (y-1, x-1) of this lane's two pixels
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int gx = tx0 + wcol + j * 4 + q;
        pyb[j] = (gy < H && gx < W) ? (float)(gy - 1) : -100000.f;   // dead pixel: sample rejected
        pxb[j] = (float)(gx - 1);
      }
      // AFF: affine parameters of this thread's two (pixel, group) items; logits stay in the staged block
      float af[2][6];
      const __nv_bfloat16* lgp[2] = {nullptr, nullptr};
      const float* sbias = reinterpret_cast<const float*>(smem + Smem::ABIAS_OFF);
      if constexpr (AFF) {
        cp_async_wait<1>();                                    // this tile's block (the next one may be in flight)
        __syncwarp();
        const int g = (l * DG) / 8;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const __nv_bfloat16* pp = reinterpret_cast<const __nv_bfloat16*>(
              smem + Smem::OFFS_OFF + (warp * NOB + (tl & 1)) * OFF_WARP_BUF + (j * 4 + q) * C::APX);
          const bool live = pyb[j] > -50000.f;                 // dead pixels were not staged
#pragma unroll
          for (int e = 0; e < 4; ++e) af[j][e] = live ? __bfloat162float(pp[g * 4 + e]) + sbias[g * 4 + e] : 0.f;
#pragma unroll
          for (int e = 0; e < 2; ++e)
            af[j][4 + e] = live ? __bfloat162float(pp[4 * DG + g * 2 + e]) + sbias[4 * DG + g * 2 + e] : 0.f;
          lgp[j] = pp + 6 * DG + g * 9;
        }
      }
      // Border tiles: the window cells outside the image are not loaded; zero them so that the gather
      // needs no per-corner validity test (an outside corner reads 0, which is DCNv2's zero padding
      // and its "p <= -1 or p >= size -> 0" rule at once).  Safe against readers of the tile that used
      // this buffer before (tl-2): no warp is more than NSA taps behind.  All 16 producer warps take
      // this branch together (the condition is CTA-uniform).
      if (wy0 < 0 || wx0 < 0 || wy0 + WH > H || wx0 + WW > W) {
        uint8_t* wb = smem + Smem::WIN_OFF + (tl & 1) * WIN_BYTES;
        for (int i = tid; i < WH * WW * 8; i += PWARPS * 32) {
          const int cell = i >> 3;
          const int cy = wy0 + cell / WW, cx = wx0 + cell % WW;
          if ((unsigned)cy >= (unsigned)H || (unsigned)cx >= (unsigned)W)
            *reinterpret_cast<uint4*>(wb + i * 16) = make_uint4(0, 0, 0, 0);
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(PWARPS * 32) : "memory");
      }
      mbar_wait(bar_winf + 8 * (tl & 1), (tl >> 1) & 1);       // this tile's window has landed
      for (int tap = 0; tap < TAPS; ++tap, ++it) {
        const float* so = offF;
        if constexpr (!AFF) {
          cp_async_wait<NOB - 2>();                            // offsets of `it` have landed
          __syncwarp();                                        // ... and everyone left buffer (it-1) % NOB
          prefetch_next();
          so = offF + oring * (OFF_WARP_BUF / 4);
          if (++oring == NOB) oring = 0;
        }
        const int ti = (tap * 11) >> 5, tj = tap - ti * 3;       // tap / 3 for tap < 9
        uint32_t res[2][4];
        // The samples of a thread (2 items x NSUB groups per 16-byte chunk) are advanced in lock step
        // (phase by phase), so that the scheduler always has independent dependency chains per warp: the
        // kernel is latency bound (ncu: "wait" and scoreboard stalls dominate with 4 warps per scheduler).
        constexpr int NI = 2 * NSUB;         // samples per thread and tap
        constexpr int CW = 4 / NSUB;         // 32-bit words per corner (8 or 4 channels)
        float wy0f[NI], wy1f[NI], wx0f[NI], wx1f[NI];
        int y0[NI], x0[NI], ry[NI], rx[NI];
        bool inwin[NI];
        bool allin = true;
#pragma unroll
        for (int u = 0; u < NI; ++u) {
          const int j = u / NSUB, sb = u % NSUB;
          const int g = NSUB == 2 ? 2 * l + sb : (l * DG) / 8;
          const int col = (j * 4 + q) ^ (((g >> 2) & 1) << 2);
          float dy, dx, mk;
          if constexpr (AFF) {                                 // models/networks.py:302-313, as affine_offsets_kernel
            const float r0 = (float)(ti - 1), r1 = (float)(tj - 1);
            dy = af[j][0] * r0 + af[j][1] * r1 - r0 + af[j][4];
            dx = af[j][2] * r0 + af[j][3] * r1 - r1 + af[j][5];
            const float lg = __bfloat162float(lgp[j][tap]) + sbias[6 * DG + g * 9 + tap];
            mk = pyb[j] > -50000.f ? 1.f / (1.f + __expf(-lg)) : 0.f;
          } else {
            dy = so[(0 * DG + g) * PLW + col];
            dx = so[(1 * DG + g) * PLW + col];
            mk = so[(2 * DG + g) * PLW + col];
          }
          if (EAVSR_ABL & 32) { dy = 0.25f; dx = 0.25f; mk = 0.5f; }
          const float py = (pyb[j] + (float)ti) + dy;
          const float px = (pxb[j] + (float)tj) + dx;
          // floor via a saturating float->int conversion (NaN -> 0, +-huge -> INT_MIN/MAX: such cells fail
          // the window test and are rejected by the validity tests of the far path)
          y0[u] = __float2int_rd(py); x0[u] = __float2int_rd(px);
          const float ly = py - (float)y0[u], lx = px - (float)x0[u];
          wy0f[u] = mk * (1.f - ly); wy1f[u] = mk * ly; wx0f[u] = 1.f - lx; wx1f[u] = lx;
          // both rows / columns of the 2x2 cell inside the window?
          ry[u] = y0[u] - wy0; rx[u] = x0[u] - wx0;
          inwin[u] = (unsigned)ry[u] < (unsigned)(WH - 1) && (unsigned)rx[u] < (unsigned)(WW - 1);
          allin = allin && inwin[u];
        }
        uint32_t v[NI][4][CW];               // [sample][corner][words]
        auto load_win = [&](int u) {
          if (EAVSR_ABL & 8) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int e = 0; e < CW; ++e) v[u][c][e] = (uint32_t)(ry[u] + c + e);
            return;
          }
          const uint32_t a00 = win + (uint32_t)(ry[u] * WW + rx[u]) * 128u + (u % NSUB) * 8u;
          const uint32_t ad[4] = {a00, a00 + 128u, a00 + WW * 128u, a00 + WW * 128u + 128u};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (CW == 4) {
              const uint4 t = lds128(ad[c]);
              v[u][c][0] = t.x; v[u][c][1] = t.y; v[u][c][CW - 2] = t.z; v[u][c][CW - 1] = t.w;
            } else {
              const uint2 t = lds64(ad[c]);
              v[u][c][0] = t.x; v[u][c][1] = t.y;
            }
          }
        };
        if (__all_sync(0xffffffffu, allin)) {                  // warp-uniform common case: no tests
#pragma unroll
          for (int u = 0; u < NI; ++u) load_win(u);
        } else {
#pragma unroll
          for (int u = 0; u < NI; ++u) {
            if (inwin[u]) {
              load_win(u);
            } else {                                           // far sample: global gather with explicit validity
              const int yy = y0[u], xx = x0[u];
              wy0f[u] = ((unsigned)yy < (unsigned)H) ? wy0f[u] : 0.f;
              wy1f[u] = ((unsigned)yy + 1u < (unsigned)H) ? wy1f[u] : 0.f;
              wx0f[u] = ((unsigned)xx < (unsigned)W) ? wx0f[u] : 0.f;
              wx1f[u] = ((unsigned)xx + 1u < (unsigned)W) ? wx1f[u] : 0.f;
              const int ys = min(max(yy, -1), H), xs = min(max(xx, -1), W);      // keep the +1 below defined
              const int cy0 = min(max(ys, 0), H - 1), cy1 = min(max(ys + 1, 0), H - 1);
              const int cx0 = min(max(xs, 0), W - 1), cx1 = min(max(xs + 1, 0), W - 1);
              const uint32_t b00 = (uint32_t)(cy0 * W + cx0) * CH + l * 8 + (u % NSUB) * 4;
              const uint32_t sx = (uint32_t)(cx1 - cx0) * CH, sy = (uint32_t)((cy1 - cy0) * W) * CH;
              const uint32_t bo[4] = {b00, (uint32_t)(b00 + sx), (uint32_t)(b00 + sy), (uint32_t)(b00 + sy + sx)};
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                if (CW == 4) {
                  const uint4 t = __ldg(reinterpret_cast<const uint4*>(xn + bo[c]));
                  v[u][c][0] = t.x; v[u][c][1] = t.y; v[u][c][CW - 2] = t.z; v[u][c][CW - 1] = t.w;
                } else {
                  const uint2 t = __ldg(reinterpret_cast<const uint2*>(xn + bo[c]));
                  v[u][c][0] = t.x; v[u][c][1] = t.y;
                }
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < NI; ++u) {
          const int j = u / NSUB, sb = u % NSUB;
          const float w00 = wy0f[u] * wx0f[u], w01 = wy0f[u] * wx1f[u], w10 = wy1f[u] * wx0f[u], w11 = wy1f[u] * wx1f[u];
          if (EAVSR_ABL & 4) {
#pragma unroll
            for (int e = 0; e < CW; ++e) res[j][sb * CW + e] = v[u][0][e] ^ v[u][1][e] ^ v[u][2][e] ^ v[u][3][e] ^ __float_as_uint(w00 + w11);
          } else if (BLEND16) {
            const uint32_t p00 = pack_bf16x2(w00, w00), p01 = pack_bf16x2(w01, w01);
            const uint32_t p10 = pack_bf16x2(w10, w10), p11 = pack_bf16x2(w11, w11);
#pragma unroll
            for (int e = 0; e < CW; ++e)
              res[j][sb * CW + e] = hfma2_bf16(p11, v[u][3][e], hfma2_bf16(p10, v[u][2][e], hfma2_bf16(p01, v[u][1][e], hmul2_bf16(p00, v[u][0][e]))));
          } else {
#pragma unroll
            for (int e = 0; e < CW; ++e) {
              const float lo = w00 * bf16lo_to_f32(v[u][0][e]) + w01 * bf16lo_to_f32(v[u][1][e]) +
                               w10 * bf16lo_to_f32(v[u][2][e]) + w11 * bf16lo_to_f32(v[u][3][e]);
              const float hi = w00 * bf16hi_to_f32(v[u][0][e]) + w01 * bf16hi_to_f32(v[u][1][e]) +
                               w10 * bf16hi_to_f32(v[u][2][e]) + w11 * bf16hi_to_f32(v[u][3][e]);
              res[j][sb * CW + e] = pack_bf16x2(lo, hi);
            }
          }
        }
        if (it >= NSA) mbar_wait(bar_empty + 8 * ring, ring_ph ^ 1);
        const uint32_t aStage = sA + ring * A_TILE;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int m = wrow * 16 + wcol + j * 4 + q;
          const uint32_t dst = aStage + sw128_offset(m, l * 16);
          if (!(EAVSR_ABL & 16) || res[j][0] == 0x12345678u)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst), "r"(res[j][0]), "r"(res[j][1]),
                       "r"(res[j][2]), "r"(res[j][3]));
        }
        if (!(EAVSR_ABL & 1)) fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_full + 8 * ring);
          if (tap == TAPS - 1) mbar_arrive(bar_wine + 8 * (tl & 1));   // this warp is done with the window
        }
        if (++ring == NSA) { ring = 0; ring_ph ^= 1; }
        if (tap == 1 && tl >= 1) epilogue(tl - 1);
      }
      if constexpr (AFF) {
        __syncwarp();                                          // every lane is done with this tile's block
        prefetch_tile(tl + 2);
      }
    }
    cp_async_wait<0>();
    epilogue(my_tiles - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PWARPS) tmem_dealloc<TMEM_COLS>(tmem_d);
}

}  // namespace win
}  // namespace eavsr
