// Fourth-generation tcgen05 DCNv2 forward (bf16 features, deform_groups = 8, W % 4 == 0): the window kernel
// (dcn_fwd_win.cuh) with the two things its timing ablations pointed at taken out of the producers'
// instruction stream:
//
//   * offsets / masks are staged by the TMA unit, not by the producer warps.  A dedicated loader warp issues three
//     5-D tensor loads per (tile, tap) -- dy, dx and mask planes of the 8 groups for the 8 x 16 tile, viewed as
//     (x, y, 2k+comp, group, n) so that the 8 group planes of one tap come in ONE box -- into a 4-stage ring, ahead
//     of the producers by whole taps; out-of-image pixels are zero-filled (mask 0 => the sample contributes 0, no
//     dead-pixel special case).  The old kernel spent 28 instructions per (warp, tap) on cp.async address arithmetic,
//     32 LDGSTS per tap on the LSU, and a wait + __syncwarp in front of every tap (ablation: -19 us without them).
//   * a lane <-> work mapping for which the NATURAL [plane][y][x] layout the TMA unit writes is bank-conflict free,
//     together with the window gather and the swizzled A-tile store: a warp owns 32 pixels (two tile rows) and two of
//     the eight "group rotations" k; lane (q, l) handles pixel 8q + r(l) of the block and, for rotation k, group
//     l ^ k.  Offsets: the 32 lanes read 32 different pixels of (possibly different) planes -- plane strides are
//     multiples of 32 words, so the bank is the pixel index: 32 distinct banks.  Gather: the 8 lanes of a quarter
//     warp read 8 different 16-byte chunks (l ^ k is a bijection in l).  A-tile store: row m, chunk g ^ (m & 7) =
//     (l ^ k) ^ r(l), a bijection in l because r(l) = alpha * l in GF(8) makes l -> l ^ r(l) one (I + R is
//     invertible).  One pixel per lane also halves the per-lane position state.
//   * the loader warp also owns the window and weight-tile copies (3-stage weight ring refilled on the MMAs' own
//     completion barrier), so the MMA thread only waits and issues.
//
// Everything else -- 8 x 16 tile + 5-pixel apron window (double buffered), far-sample global fallback, bf16x2 HFMA2
// or fp32 blend, two TMEM accumulators, epilogue by the producers one tile behind -- is the window kernel's.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "dcn_fwd_win.cuh"

namespace eavsr {
namespace win2 {

using win::hfma2_bf16;
using win::hmul2_bf16;
using win::lds128;

constexpr int PWARPS = 16;
constexpr int THREADS = (PWARPS + 2) * 32;   // 16 producers, MMA issuer, loader
constexpr int TH = 8, TW = 16, CH = 64, TAPS = 9, DG = 8;
constexpr int PAD = 5, WH = TH + 2 * PAD, WW = TW + 2 * PAD;   // 18 x 26 window
constexpr int WIN_BYTES = WH * WW * 128;     // 59 904
constexpr int A_TILE = 128 * CH * 2;         // 16 KB
constexpr int B_TILE = CH * CH * 2;          // 8 KB
constexpr int NSA = 2, NSB = 3, NOS = 4;
constexpr int O_PLANE = TH * TW * 4;         // 512 B: one (component, group) plane of the tile
constexpr int O_COMP = DG * O_PLANE;         // 4 KB: the 8 group planes of one component
constexpr int O_STAGE = 3 * O_COMP;          // 12 KB: dy | dx | mask
constexpr int TMEM_COLS = 128;

struct Smem {
  static constexpr int WIN_OFF = 0;
  static constexpr int A_OFF = WIN_OFF + 2 * WIN_BYTES;            // 119 808 (1024-aligned)
  static constexpr int B_OFF = A_OFF + NSA * A_TILE;
  static constexpr int O_OFF = B_OFF + NSB * B_TILE;               // 128-byte aligned TMA destinations
  static constexpr int BAR_OFF = O_OFF + NOS * O_STAGE;
  // afull[NSA] aempty[NSA] bfull[NSB] bempty[NSB] ofull[NOS] oempty[NOS] accf[2] acce[2] winf[2] wine[2]
  static constexpr int NBARS = 2 * NSA + 2 * NSB + 2 * NOS + 8;
  static constexpr int TOTAL = BAR_OFF + NBARS * 8 + 16;
  static constexpr int DYN = TOTAL + 1024;
  static_assert(A_OFF % 1024 == 0 && B_OFF % 1024 == 0 && O_OFF % 128 == 0, "operand / TMA alignment");
  static_assert(DYN <= 232448, "shared memory budget");
};

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::
          "r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
      : "memory");
}

template <bool BLEND16>
__global__ void __launch_bounds__(THREADS, 1)
dcn_fwd_win2_kernel(const __nv_bfloat16* __restrict__ x, const __grid_constant__ CUtensorMap tmOff,
                    const __grid_constant__ CUtensorMap tmMask, const uint8_t* __restrict__ wpacked,
                    const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W, long long xs_n,
                    long long os_n, int tiles_x, int tiles_per_img, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  asm volatile("mov.u32 %0, %0;\n" : "+r"(sbase));      // opaque: keep it in a register instead of re-deriving it per use
  uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t sWin = sbase + Smem::WIN_OFF, sA = sbase + Smem::A_OFF, sB = sbase + Smem::B_OFF, sO = sbase + Smem::O_OFF;
  const uint32_t bars = sbase + Smem::BAR_OFF;
  const uint32_t bar_afull = bars, bar_aempty = bar_afull + NSA * 8, bar_bfull = bar_aempty + NSA * 8;
  const uint32_t bar_bempty = bar_bfull + NSB * 8, bar_ofull = bar_bempty + NSB * 8, bar_oempty = bar_ofull + NOS * 8;
  const uint32_t bar_accf = bar_oempty + NOS * 8, bar_acce = bar_accf + 16, bar_winf = bar_acce + 16, bar_wine = bar_winf + 16;
  const uint32_t tmem_slot_addr = bar_wine + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Smem::BAR_OFF + Smem::NBARS * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NSA; ++s) { mbar_init(bar_afull + 8 * s, PWARPS); mbar_init(bar_aempty + 8 * s, 1); }
    for (int s = 0; s < NSB; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
    for (int s = 0; s < NOS; ++s) { mbar_init(bar_ofull + 8 * s, 1); mbar_init(bar_oempty + 8 * s, PWARPS); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accf + 8 * b, 1);
      mbar_init(bar_acce + 8 * b, PWARPS);
      mbar_init(bar_winf + 8 * b, 1);
      mbar_init(bar_wine + 8 * b, PWARPS);
    }
    fence_mbar_init();
  }
  // window cells outside the image are never loaded; zero both buffers once so that whatever they hold later is finite
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
  auto tile_coords = [&](int tl, int& n, int& ty0, int& tx0) {
    const int tile = first + tl * (int)gridDim.x;
    n = tile / tiles_per_img;
    const int rem = tile - n * tiles_per_img;
    ty0 = (rem / tiles_x) * TH;
    tx0 = (rem % tiles_x) * TW;
  };

  if (warp == PWARPS + 1) {
    // ============ loader: offsets / masks (TMA tensor loads), window rows and weight tiles (bulk copies) ============
    // Step j serves tap j of this CTA's stream.  Every wait is on work of an EARLIER step whose inputs were issued at
    // an earlier step, so the loader can run ahead (up to NOS taps) without ever deadlocking the pipeline.
    if (elect_one()) {
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
      int tl = 0, tap = 0, n = 0, ty0 = 0, tx0 = 0;
      load_window(0);
      for (int j = 0; j < n_iters; ++j) {
        if (tap == 0) tile_coords(tl, n, ty0, tx0);
        if (tap == NOS && tl + 1 < my_tiles) {
          // window of tile tl+1: its buffer was last read by tile tl-1, which every producer has left (the loader
          // is past tap NOS of tile tl, i.e. the offsets stage of tap 9*tl - 1 has been released by all of them)
          if (tl >= 1) mbar_wait(bar_wine + 8 * ((tl + 1) & 1), (((tl + 1) >> 1) - 1) & 1);
          load_window(tl + 1);
        }
        {  // offsets / masks of tap j
          const int s = j % NOS;
          if (j >= NOS) mbar_wait(bar_oempty + 8 * s, ((j / NOS) - 1) & 1);
          const uint32_t bar = bar_ofull + 8 * s, dst = sO + s * O_STAGE;
          mbar_arrive_expect_tx(bar, O_STAGE);
          tma_load_5d(dst, &tmOff, tx0, ty0, 2 * tap, 0, n, bar);
          tma_load_5d(dst + O_COMP, &tmOff, tx0, ty0, 2 * tap + 1, 0, n, bar);
          tma_load_5d(dst + 2 * O_COMP, &tmMask, tx0, ty0, tap, 0, n, bar);
        }
        {  // weight tile of tap j
          const int s = j % NSB;
          if (j >= NSB) mbar_wait(bar_bempty + 8 * s, ((j / NSB) - 1) & 1);
          mbar_arrive_expect_tx(bar_bfull + 8 * s, B_TILE);
          bulk_g2s(sB + s * B_TILE, wpacked + (size_t)tap * B_TILE, B_TILE, bar_bfull + 8 * s);
        }
        if (++tap == TAPS) { tap = 0; ++tl; }
      }
    }
    __syncwarp();
  } else if (warp == PWARPS) {
    // ============ MMA issuer ============
    if (elect_one()) {
      constexpr uint32_t IDESC = umma_idesc_bf16(128, CH);
      const uint64_t a_base = umma_desc_sw128_kmajor(sA), b_base = umma_desc_sw128_kmajor(sB);
      int tap = 0, tl = 0;
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % NSA, sb = it % NSB, buf = tl & 1;
        if (tap == 0 && tl >= 2) mbar_wait(bar_acce + 8 * buf, ((tl >> 1) - 1) & 1);
        mbar_wait(bar_afull + 8 * s, (it / NSA) & 1);
        mbar_wait(bar_bfull + 8 * sb, (it / NSB) & 1);
        tc_fence_after();
        const uint64_t a_d = a_base + (uint64_t)((s * A_TILE) >> 4), b_d = b_base + (uint64_t)((sb * B_TILE) >> 4);
        const uint32_t d = tmem_d + buf * CH;
#pragma unroll
        for (int k = 0; k < CH / 16; ++k) umma_bf16(d, a_d + 2 * k, b_d + 2 * k, IDESC, (tap | k) != 0);
        umma_commit(bar_aempty + 8 * s);
        umma_commit(bar_bempty + 8 * sb);
        if (tap == TAPS - 1) umma_commit(bar_accf + 8 * buf);
        if (++tap == TAPS) { tap = 0; ++tl; }
      }
    }
    __syncwarp();
  } else {
    // ============ producers: window gather -> blend -> swizzled A stage ============
    const int q = lane >> 3, l = lane & 7;
    const int rl = ((l << 1) ^ ((l & 4) ? 0xB : 0)) & 7;       // alpha * l in GF(8) = {0,2,4,6,3,1,7,5}
    const int blk = warp & 3, kp = warp >> 2;                  // 32-pixel block, pair of group rotations
    const int p = q * 8 + rl;                                  // pixel inside the block
    const int m = blk * 32 + p;                                // tile pixel = A-tile row = TMEM lane
    const int trow = m >> 4, tcol = m & 15;
    int gsel[2];                                               // this lane's group for the two samples of a tap
    uint32_t ooff[2], soff[2];                                 // byte offsets: offset plane word / A-tile chunk
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      gsel[u] = l ^ (2 * kp + u);
      ooff[u] = (uint32_t)(gsel[u] * O_PLANE + m * 4);
      soff[u] = sw128_offset((uint32_t)m, (uint32_t)gsel[u] * 16u);
    }

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
      const int mm = qd * 32 + lane;
      const int gy = ty0 + (mm >> 4), gx = tx0 + (mm & 15);
      if (gy < H && gx < W) {
        __nv_bfloat16* op = out + (size_t)n * os_n + ((size_t)gy * W + gx) * CH + cq * 16;
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

    int it = 0, ring = 0, oring = 0;
    uint32_t ring_ph = 0, oring_ph = 0;
    for (int tl = 0; tl < my_tiles; ++tl) {
      int n, ty0, tx0;
      tile_coords(tl, n, ty0, tx0);
      const __nv_bfloat16* xn = x + (size_t)n * xs_n;
      const int wy0 = ty0 - PAD, wx0 = tx0 - PAD;
      const uint32_t winb = sWin + (tl & 1) * WIN_BYTES;
      const float pyb = (float)(ty0 + trow - 1), pxb = (float)(tx0 + tcol - 1);     // (y - 1, x - 1) of this lane's pixel
      // Border tiles: zero the window cells outside the image (they are not loaded), so that the gather needs no
      // per-corner validity test.  CTA-uniform condition: all 16 producer warps take the branch together.
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
      float tif = 0.f, tjf = 0.f;                              // tap row / column as floats (no division per tap)
      for (int tap = 0; tap < TAPS; ++tap, ++it) {
        mbar_wait(bar_ofull + 8 * oring, oring_ph);            // offsets / masks of this tap have landed
        const uint32_t ob = sO + oring * O_STAGE;
        float wy0f[2], wy1f[2], wx0f[2], wx1f[2];
        int y0[2], x0[2], ry[2], rx[2];
        bool inwin[2];
        bool allin = true;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float dy, dx, mk;
          asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(dy) : "r"(ob + ooff[u]));
          asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(dx) : "r"(ob + ooff[u] + O_COMP));
          asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(mk) : "r"(ob + ooff[u] + 2 * O_COMP));
          const float py = (pyb + tif) + dy, px = (pxb + tjf) + dx;
          // floor via a saturating float->int conversion (NaN -> 0, +-huge -> INT_MIN/MAX: such cells fail the window
          // test and are rejected by the validity tests of the far path)
          y0[u] = __float2int_rd(py); x0[u] = __float2int_rd(px);
          const float ly = py - (float)y0[u], lx = px - (float)x0[u];
          wy0f[u] = mk * (1.f - ly); wy1f[u] = mk * ly; wx0f[u] = 1.f - lx; wx1f[u] = lx;
          ry[u] = y0[u] - wy0; rx[u] = x0[u] - wx0;
          inwin[u] = (unsigned)ry[u] < (unsigned)(WH - 1) && (unsigned)rx[u] < (unsigned)(WW - 1);
          allin = allin && inwin[u];
        }
        uint32_t v[2][4][4];
        auto load_win = [&](int u) {
          const uint32_t a00 = winb + (uint32_t)(ry[u] * WW + rx[u]) * 128u + (uint32_t)gsel[u] * 16u;
          const uint32_t ad[4] = {a00, a00 + 128u, a00 + WW * 128u, a00 + WW * 128u + 128u};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 t = lds128(ad[c]);
            v[u][c][0] = t.x; v[u][c][1] = t.y; v[u][c][2] = t.z; v[u][c][3] = t.w;
          }
        };
        if (__all_sync(0xffffffffu, allin)) {                  // warp-uniform common case: no tests
          load_win(0);
          load_win(1);
        } else {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
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
              const uint32_t b00 = (uint32_t)(cy0 * W + cx0) * CH + gsel[u] * 8;
              const uint32_t sx = (uint32_t)(cx1 - cx0) * CH, sy = (uint32_t)((cy1 - cy0) * W) * CH;
              const uint32_t bo[4] = {b00, (uint32_t)(b00 + sx), (uint32_t)(b00 + sy), (uint32_t)(b00 + sy + sx)};
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint4 t = __ldg(reinterpret_cast<const uint4*>(xn + bo[c]));
                v[u][c][0] = t.x; v[u][c][1] = t.y; v[u][c][2] = t.z; v[u][c][3] = t.w;
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_oempty + 8 * oring);    // this warp has read its offsets: the stage may refill
        if (++oring == NOS) { oring = 0; oring_ph ^= 1; }
        uint32_t res[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float w00 = wy0f[u] * wx0f[u], w01 = wy0f[u] * wx1f[u], w10 = wy1f[u] * wx0f[u], w11 = wy1f[u] * wx1f[u];
          if (BLEND16) {
            const uint32_t p00 = pack_bf16x2(w00, w00), p01 = pack_bf16x2(w01, w01);
            const uint32_t p10 = pack_bf16x2(w10, w10), p11 = pack_bf16x2(w11, w11);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              res[u][e] = hfma2_bf16(p11, v[u][3][e], hfma2_bf16(p10, v[u][2][e], hfma2_bf16(p01, v[u][1][e], hmul2_bf16(p00, v[u][0][e]))));
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float lo = w00 * bf16lo_to_f32(v[u][0][e]) + w01 * bf16lo_to_f32(v[u][1][e]) +
                               w10 * bf16lo_to_f32(v[u][2][e]) + w11 * bf16lo_to_f32(v[u][3][e]);
              const float hi = w00 * bf16hi_to_f32(v[u][0][e]) + w01 * bf16hi_to_f32(v[u][1][e]) +
                               w10 * bf16hi_to_f32(v[u][2][e]) + w11 * bf16hi_to_f32(v[u][3][e]);
              res[u][e] = pack_bf16x2(lo, hi);
            }
          }
        }
        if (it >= NSA) mbar_wait(bar_aempty + 8 * ring, ring_ph ^ 1);
        const uint32_t aStage = sA + ring * A_TILE;
#pragma unroll
        for (int u = 0; u < 2; ++u)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(aStage + soff[u]), "r"(res[u][0]), "r"(res[u][1]),
                       "r"(res[u][2]), "r"(res[u][3]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_afull + 8 * ring);
          if (tap == TAPS - 1) mbar_arrive(bar_wine + 8 * (tl & 1));   // this warp is done with the window
        }
        if (++ring == NSA) { ring = 0; ring_ph ^= 1; }
        tjf += 1.f;
        if (tjf == 3.f) { tjf = 0.f; tif += 1.f; }
        if (tap == 1 && tl >= 1) epilogue(tl - 1);
      }
    }
    epilogue(my_tiles - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PWARPS) tmem_dealloc<TMEM_COLS>(tmem_d);
}

}  // namespace win2
}  // namespace eavsr
