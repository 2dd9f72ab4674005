// Fifth-generation tcgen05 DCNv2 forward (bf16 features, deform_groups = 8, W % 4 == 0): the sampled column tile never
// touches shared memory.  The ncu capture of the fourth generation (profiles/r2_dcn_fwd_win2_ncu.txt) shows a kernel
// that is neither issue- (54 %), shared-memory- (36 %) nor HBM-bound (23 %): its 16 producer warps all feed the SAME
// 16 KB A stage of a 2-deep ring, so every tap is a 16-warp rendezvous (42 % of the "stage free" waits take the slow
// path, and whenever one warp has a far sample -- 90 % of the taps at sigma = 2 -- the other fifteen wait for its
// global loads).  Shared memory has no room for a deeper ring; tensor memory has 384 idle columns.  So:
//
//   * A operand in TMEM ("TS" form of tcgen05.mma: D[tmem] += A[tmem] * B[smem]).  One tap's 128 x 64 bf16 tile is
//     32 columns; the ring is NSA = 12 taps deep next to the two 64-column accumulators (512 columns in all).  A
//     producer thread owns ONE pixel (TMEM lane = tile pixel = its lane in the warp's lane quadrant) and, for a tap,
//     all 8 deformable groups of it, and writes the pixel's 64 channels with a single tcgen05.st.32x32b.x32.  No
//     swizzled st.shared, no fence.proxy.async, 32 KB of shared memory back.
//   * a tap is produced by FOUR warps (one per lane quadrant), not sixteen: the four warp sets work on four
//     consecutive taps of the CTA's (tile, tap) stream at once and may drift up to NSA taps apart, so a set that hits
//     far samples no longer stalls the other twelve warps.  The MMA thread consumes the ring in stream order.
//   * bank-conflict-free gathers need the 8 lanes of a quarter warp to read 8 different 16-byte chunks, while
//     tcgen05.st wants every lane to present the same columns.  Lane l therefore samples group l ^ u at step u
//     (u = 0..7, a Latin square).  Bit 0 of that permutation is undone in registers (one stage of an XOR butterfly of
//     selects); bits 1-2 move whole 16-channel K blocks and are undone by the MMA issuer: rows are in four classes
//     c = (pixel >> 1) & 3, a class-c row keeps group pair j at column block j ^ c, and there is one MMA per (j, c)
//     with the other classes' accumulator rows switched off by the disable-output-lane mask of tcgen05.mma.
//   * the feature window (6-pixel apron) comes in ONE 4-D tensor load per tile with out-of-image cells zero-filled by
//     the TMA unit (no border pass, no producer-wide barrier); offsets / masks as in the fourth generation (three 5-D
//     tensor loads per tap, ring of NOS = 5 taps) by a loader warp that waits for nothing but its own ring, and a
//     producer reads a tap's 24 values into registers first and hands the stage straight back (pure prefetch ring).
//     The weight tiles (ring of 2) and the windows come from a second loader warp that follows the MMAs.
//   * mbarrier waits park in hardware (try_wait with a suspend-time hint); tile coordinates without integer division.
//
// Arithmetic (bilinear weights, bf16x2 or fp32 blend, accumulation order over taps and channels) is the fourth
// generation's, so the two kernels agree bit for bit (tests/test_gpu_ops.py asserts it).  DESIGN.md 3.1c has the
// step-by-step timings (68.1 -> 46.9 us) and the ablations.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "dcn_fwd_win.cuh"
#include "dcn_fwd_win2.cuh"

namespace eavsr {
namespace win3 {

using win::hfma2_bf16;
using win::hmul2_bf16;
using win::lds128;
using win2::tma_load_5d;

constexpr int PWARPS = 16, SETS = 4;
constexpr int THREADS = (PWARPS + 3) * 32;     // 16 producers, MMA issuer, offsets loader, window / weights loader
constexpr int TH = 8, TW = 16, CH = 64, TAPS = 9, DG = 8;
#ifndef EAVSR_WIN3_PAD
#define EAVSR_WIN3_PAD 6
#endif
constexpr int PAD = EAVSR_WIN3_PAD, WH = TH + 2 * PAD, WW = TW + 2 * PAD;   // 20 x 28 window
constexpr int WIN_BYTES = WH * WW * 128;       // 71 680
constexpr int B_TILE = CH * CH * 2;            // 8 KB
#ifndef EAVSR_WIN3_NSA
#define EAVSR_WIN3_NSA 12
#endif
#ifndef EAVSR_WIN3_NSB
#define EAVSR_WIN3_NSB 2
#endif
#ifndef EAVSR_WIN3_NOS
#define EAVSR_WIN3_NOS 5
#endif
#ifndef EAVSR_ABL3
#define EAVSR_ABL3 0      // timing ablations (tools/abl_build.sh): wrong results on purpose, never set in the product build
#endif
constexpr int NSA = EAVSR_WIN3_NSA, NSB = EAVSR_WIN3_NSB, NOS = EAVSR_WIN3_NOS;
constexpr int O_PLANE = TH * TW * 4;           // 512 B: one (component, group) plane of the tile
constexpr int O_COMP = DG * O_PLANE;           // 4 KB: the 8 group planes of one component
constexpr int O_STAGE = 3 * O_COMP;            // 12 KB: dy | dx | mask
constexpr int A_COLS = CH / 2;                 // 32 TMEM columns per tap (two bf16 per column)
constexpr int TMEM_COLS = 512;
constexpr int A_COL0 = 2 * CH;                 // behind the two accumulators
static_assert(NSA >= SETS && A_COL0 + NSA * A_COLS <= TMEM_COLS, "TMEM budget");
static_assert(NOS >= SETS && NOS < 2 * SETS && NSA < 4 * SETS, "ring counters wrap at most once per step");

struct Smem {
  static constexpr int WIN_OFF = 0;
  static constexpr int O_OFF = (WIN_OFF + 2 * WIN_BYTES + 4095) / 4096 * 4096;   // 4 KB aligned: plane index ORs in
  static constexpr int B_OFF = O_OFF + NOS * O_STAGE;
  static constexpr int BAR_OFF = B_OFF + NSB * B_TILE;
  static constexpr int NBARS = 2 * NSA + 2 * NSB + 2 * NOS + 8;
  static constexpr int TOTAL = BAR_OFF + NBARS * 8 + 16;
  static constexpr int DYN = TOTAL + 4096;
  static_assert(B_OFF % 1024 == 0 && O_OFF % 4096 == 0 && O_STAGE % 4096 == 0, "operand / TMA alignment");
  static_assert(DYN <= 232448, "shared memory budget");
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
          "r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A is 128 lanes x (K / 2) columns, row m in lane m, elements 2j, 2j + 1 of the
// row in column j (low half = even element).
// `off_mask`: bit i of the byte = 1 keeps accumulator rows 8j + i (all j) untouched (disable-output-lane vector).
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate, uint32_t off_mask) {
  const uint32_t mk = off_mask * 0x01010101u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(mk)
      : "memory");
}
// thread i of the warp writes 32 consecutive columns of TMEM lane (quadrant base + i)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// mbarrier wait that parks the thread in hardware (try_wait with a suspend-time hint) instead of polling with
// nanosleep: no issue slots burnt while waiting, and the wake-up is not quantised to the sleep period.  Bounded
// like mbar_wait (a protocol bug must trap, not hang).
__device__ __forceinline__ void mbar_wait_park(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  if (done) return;
  const long long t0 = clock64();
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(1000000u)
        : "memory");
    if (!done && clock64() - t0 > 4000000000ll) __trap();
  } while (!done);
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

template <bool BLEND16>
__global__ void __launch_bounds__(THREADS, 1)
dcn_fwd_win3_kernel(const __nv_bfloat16* __restrict__ x, const __grid_constant__ CUtensorMap tmX,
                    const __grid_constant__ CUtensorMap tmOff, const __grid_constant__ CUtensorMap tmMask,
                    const uint8_t* __restrict__ wpacked, const __nv_bfloat16* __restrict__ bias,
                    __nv_bfloat16* __restrict__ out, int H, int W, long long xs_n, long long os_n, int tiles_x,
                    int tiles_per_img, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint32_t sbase = (smem_u32(smem_raw) + 4095u) & ~4095u;
  asm volatile("mov.u32 %0, %0;\n" : "+r"(sbase));      // opaque: keep it in a register instead of re-deriving it per use
  uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t sWin = sbase + Smem::WIN_OFF, sB = sbase + Smem::B_OFF, sO = sbase + Smem::O_OFF;
  const uint32_t bars = sbase + Smem::BAR_OFF;
  const uint32_t bar_afull = bars, bar_aempty = bar_afull + NSA * 8, bar_bfull = bar_aempty + NSA * 8;
  const uint32_t bar_bempty = bar_bfull + NSB * 8, bar_ofull = bar_bempty + NSB * 8, bar_oempty = bar_ofull + NOS * 8;
  const uint32_t bar_accf = bar_oempty + NOS * 8, bar_acce = bar_accf + 16, bar_winf = bar_acce + 16, bar_wine = bar_winf + 16;
  const uint32_t tmem_slot_addr = bar_wine + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Smem::BAR_OFF + Smem::NBARS * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NSA; ++s) { mbar_init(bar_afull + 8 * s, PWARPS / SETS); mbar_init(bar_aempty + 8 * s, 1); }
    for (int s = 0; s < NSB; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
    for (int s = 0; s < NOS; ++s) { mbar_init(bar_ofull + 8 * s, 1); mbar_init(bar_oempty + 8 * s, PWARPS / SETS); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accf + 8 * b, 1);
      mbar_init(bar_acce + 8 * b, PWARPS);
      mbar_init(bar_winf + 8 * b, 1);
      mbar_init(bar_wine + 8 * b, PWARPS);
    }
    fence_mbar_init();
  }
  if (warp == PWARPS) tmem_alloc<TMEM_COLS>(tmem_slot_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  const int first = blockIdx.x;
  const int my_tiles = (total_tiles - first + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_iters = my_tiles * TAPS;
  // tile -> (image, tile row, tile column) without integer division (it ran once per warp and tile, ~50 instructions
  // each): float reciprocal + one-step correction, exact for tile counts below 2^22
  const float rcp_tpi = 1.f / (float)tiles_per_img, rcp_tx = 1.f / (float)tiles_x;
  auto divmod = [](int a, int d, float rcp, int& q, int& r) {
    q = __float2int_rz((float)a * rcp);
    r = a - q * d;
    if (r < 0) { r += d; --q; }
    if (r >= d) { r -= d; ++q; }
  };
  auto tile_coords = [&](int tl, int& n, int& ty0, int& tx0) {
    const int tile = first + tl * (int)gridDim.x;
    int rem, ty, tx;
    divmod(tile, tiles_per_img, rcp_tpi, n, rem);
    divmod(rem, tiles_x, rcp_tx, ty, tx);
    ty0 = ty * TH;
    tx0 = tx * TW;
  };

  if (warp == PWARPS + 1) {
    // ============ offsets loader: three 5-D tensor loads per (tile, tap), NOS taps ahead of the producers ============
    // It waits for nothing but its own ring (a stage is released by the four warps that sampled from it), so the
    // distance it keeps ahead of the producers does not depend on the MMAs' progress.
    if (elect_one()) {
      int tl = 0, tap = 0, n = 0, ty0 = 0, tx0 = 0;
      for (int j = 0; j < n_iters; ++j) {
        if (tap == 0) tile_coords(tl, n, ty0, tx0);
        const int s = j % NOS;
        if (j >= NOS) mbar_wait_park(bar_oempty + 8 * s, ((j / NOS) - 1) & 1);
        const uint32_t bar = bar_ofull + 8 * s, dst = sO + s * O_STAGE;
        if ((EAVSR_ABL3 & 1) && j >= NOS) {
          mbar_arrive(bar);
        } else {
          mbar_arrive_expect_tx(bar, O_STAGE);
          tma_load_5d(dst, &tmOff, tx0, ty0, 2 * tap, 0, n, bar);
          tma_load_5d(dst + O_COMP, &tmOff, tx0, ty0, 2 * tap + 1, 0, n, bar);
          tma_load_5d(dst + 2 * O_COMP, &tmMask, tx0, ty0, tap, 0, n, bar);
        }
        if (++tap == TAPS) { tap = 0; ++tl; }
      }
    }
    __syncwarp();
  } else if (warp == PWARPS + 2) {
    // ============ window / weights loader: follows the MMAs (a weight slot is refilled when its MMAs complete) ============
    if (elect_one()) {
      auto load_window = [&](int tl) {
        int n, ty0, tx0;
        tile_coords(tl, n, ty0, tx0);
        const uint32_t bar = bar_winf + 8 * (tl & 1);
        if ((EAVSR_ABL3 & 2) && tl >= 2) { mbar_arrive(bar); return; }
        mbar_arrive_expect_tx(bar, WIN_BYTES);
        tma_load_4d(sWin + (tl & 1) * WIN_BYTES, &tmX, 0, tx0 - PAD, ty0 - PAD, n, bar);   // outside the image: zeros
      };
      int tl = 0, tap = 0;
      load_window(0);
      if (my_tiles > 1) load_window(1);
      for (int j = 0; j < n_iters; ++j) {
        const int s = j % NSB;
        if (j >= NSB) mbar_wait_park(bar_bempty + 8 * s, ((j / NSB) - 1) & 1);
        if ((EAVSR_ABL3 & 64) && j >= NSB) {
          mbar_arrive(bar_bfull + 8 * s);
        } else {
          mbar_arrive_expect_tx(bar_bfull + 8 * s, B_TILE);
          bulk_g2s(sB + s * B_TILE, wpacked + (size_t)tap * B_TILE, B_TILE, bar_bfull + 8 * s);
        }
        if (tap == NSB && tl >= 1 && tl + 1 < my_tiles) {
          // window of tile tl+1 into the buffer tile tl-1 used.  This step runs after the MMAs of tap (tl-1, 8)
          // completed, i.e. after every producer's last gather of tile tl-1; the barrier makes that explicit.
          mbar_wait_park(bar_wine + 8 * ((tl + 1) & 1), (((tl + 1) >> 1) - 1) & 1);
          load_window(tl + 1);
        }
        if (++tap == TAPS) { tap = 0; ++tl; }
      }
    }
    __syncwarp();
  } else if (warp == PWARPS) {
    // ============ MMA issuer: consumes the TMEM ring in stream order ============
    if (elect_one()) {
      constexpr uint32_t IDESC = umma_idesc_bf16(128, CH);
      const uint64_t b_base = umma_desc_sw128_kmajor(sB);
      int tap = 0, tl = 0;
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % NSA, sb = it % NSB, buf = tl & 1;
        if (tap == 0 && tl >= 2) mbar_wait_park(bar_acce + 8 * buf, ((tl >> 1) - 1) & 1);
        mbar_wait_park(bar_bfull + 8 * sb, (it / NSB) & 1);
        mbar_wait_park(bar_afull + 8 * s, (it / NSA) & 1);
        tc_fence_after();
        const uint64_t b_d = b_base + (uint64_t)((sb * B_TILE) >> 4);
        const uint32_t d = tmem_d + buf * CH, a = tmem_d + A_COL0 + s * A_COLS;
        // Rows are in four classes c = (pixel >> 1) & 3: a class-c row keeps the group pair j at column block j ^ c
        // (the producers only undo bit 0 of their sampling order, see below).  One MMA per (group pair j, class c)
        // with the other classes' accumulator rows masked off: 4x the tensor work on a pipe that was 7 % busy, for
        // 64 fewer selects per pixel and tap in the producers.  Every row still accumulates j = 0..3 in order.
#pragma unroll
        for (int j = 0; j < CH / 16; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c)
            umma_bf16_ts(d, a + 8 * (j ^ c), b_d + 2 * j, IDESC, (tap | j) != 0, 0xFFu ^ (3u << (2 * c)));
        umma_commit(bar_aempty + 8 * s);
        umma_commit(bar_bempty + 8 * sb);
        if (tap == TAPS - 1) umma_commit(bar_accf + 8 * buf);
        if (++tap == TAPS) { tap = 0; ++tl; }
      }
    }
    __syncwarp();
  } else {
    // ============ producers: window gather -> blend -> TMEM ============
    const int l = lane & 7;
    const int qd = warp & 3, set = warp >> 2;                  // TMEM lane quadrant (32-pixel block), warp set
    const int m = qd * 32 + lane;                              // tile pixel = TMEM lane
    const int trow = m >> 4, tcol = m & 15;
    const uint32_t lsel16 = (uint32_t)l * 16u, lsel512 = (uint32_t)l * O_PLANE;

    auto epilogue = [&](int tl) {
      const int buf = tl & 1;
      mbar_wait_park(bar_accf + 8 * buf, (tl >> 1) & 1);
      tc_fence_after();
      const int cq = set;                                      // 16-column quarter
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
      const int gy = ty0 + trow, gx = tx0 + tcol;
      if (gy < H && gx < W) {
        __nv_bfloat16* op = out + (size_t)n * os_n + ((size_t)gy * W + gx) * CH + cq * 16;
        float f[16];
        uint32_t bw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (bias) {
          const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(bias + cq * 16));
          const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(bias + cq * 16) + 1);
          bw[0] = b0.x; bw[1] = b0.y; bw[2] = b0.z; bw[3] = b0.w; bw[4] = b1.x; bw[5] = b1.y; bw[6] = b1.z; bw[7] = b1.w;
        }
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          f[e] = __uint_as_float(acc[e]) + bf16lo_to_f32(bw[e >> 1]);
          f[e + 1] = __uint_as_float(acc[e + 1]) + bf16hi_to_f32(bw[e >> 1]);
        }
#pragma unroll
        for (int e = 0; e < 16; e += 8) {
          uint4 u;
          u.x = pack_bf16x2(f[e], f[e + 1]); u.y = pack_bf16x2(f[e + 2], f[e + 3]);
          u.z = pack_bf16x2(f[e + 4], f[e + 5]); u.w = pack_bf16x2(f[e + 6], f[e + 7]);
          *reinterpret_cast<uint4*>(op + e) = u;
        }
      }
    };

    int tl = 0, tap = set, os = set, as = set, ep_done = 0, taps_in_tile = 0;
    uint32_t oph = 0, aph = 0;
    bool newtile = true;
    const __nv_bfloat16* xn = x;
    uint32_t winl = 0;                                         // window base | this lane's chunk rotation
    int wy0 = 0, wx0 = 0;
    float pyb = 0.f, pxb = 0.f;
    for (int it = set; it < n_iters; it += SETS) {
      if (newtile) {
        int n, ty0, tx0;
        tile_coords(tl, n, ty0, tx0);
        xn = x + (size_t)n * xs_n;
        wy0 = ty0 - PAD; wx0 = tx0 - PAD;
        winl = (sWin + (tl & 1) * WIN_BYTES) | lsel16;         // window buffers are 128-byte aligned
        pyb = (float)(ty0 + trow - 1); pxb = (float)(tx0 + tcol - 1);     // (y - 1, x - 1) of this lane's pixel
        mbar_wait_park(bar_winf + 8 * (tl & 1), (tl >> 1) & 1);     // this tile's window has landed
        taps_in_tile = 0;
        newtile = false;
      }
      const int ti = (tap * 11) >> 5, tj = tap - 3 * ti;
      const float pyt = pyb + (float)ti, pxt = pxb + (float)tj;
      mbar_wait_park(bar_ofull + 8 * os, oph);                      // offsets / masks of this tap have landed
      const uint32_t obl = (sO + os * O_STAGE + (uint32_t)m * 4u) | lsel512;   // stage is 4 KB aligned, m * 4 < 512

      // the tap's 8 x (dy, dx, mask) into registers first (they die one sample at a time while res[] fills up, so
      // the peak register count does not move) and the stage goes straight back to the loader: the offsets ring is
      // then NOS - 1 taps of pure prefetch instead of being held by the four warp sets for the length of a tap
      float ody[DG], odx[DG], omk[DG];
#pragma unroll
      for (int u = 0; u < DG; ++u) {
        const uint32_t oa = obl ^ (uint32_t)(u * O_PLANE);
        if (EAVSR_ABL3 & 32) {
          ody[u] = __uint_as_float(oa) * 1e-9f; odx[u] = ody[u] + 0.25f; omk[u] = 0.5f;
        } else {
          asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(ody[u]) : "r"(oa));
          asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(odx[u]) : "r"(oa + O_COMP));
          asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(omk[u]) : "r"(oa + 2 * O_COMP));
        }
      }
      // mbarrier.arrive is a release at CTA scope and __syncwarp orders the other lanes' loads before lane 0's arrive
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_oempty + 8 * os);

      uint32_t v[2][4][4];                                     // corner chunks of the two samples in flight
      uint32_t wq[2][4];                                       // their bilinear weights (bf16x2 pairs or fp32 bits)
      uint32_t res[8][4];                                      // res[u]: 8 channels of group l ^ u
      // coordinates + gather issue of sample u (group l ^ u)
      auto issue = [&](const int u) {
        const int b = u & 1;
        const float dy = ody[u], dx = odx[u], mk = omk[u];
        const float py = pyt + dy, px = pxt + dx;
        // floor via a saturating float->int conversion (NaN -> 0, +-huge -> INT_MIN/MAX: such cells fail the window
        // test and are rejected by the validity tests of the far path)
        const int y0 = __float2int_rd(py), x0 = __float2int_rd(px);
        const float ly = py - (float)y0, lx = px - (float)x0;
        float wy0f = mk * (1.f - ly), wy1f = mk * ly, wx0f = 1.f - lx, wx1f = lx;
        int ry = y0 - wy0, rx = x0 - wx0;
        if (EAVSR_ABL3 & 16) { ry = min(max(ry, 0), WH - 2); rx = min(max(rx, 0), WW - 2); }
        const bool inwin = (unsigned)ry < (unsigned)(WH - 1) && (unsigned)rx < (unsigned)(WW - 1);
        if (__all_sync(0xffffffffu, inwin) || inwin) {         // warp-uniform common case first: no divergence
          const uint32_t a00 = ((uint32_t)(ry * WW + rx) * 128u + winl) ^ (uint32_t)(u * 16);
          if (EAVSR_ABL3 & 4) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int e = 0; e < 4; ++e) v[b][c][e] = a00 + c * 4 + e;
          } else {
          const uint4 t0 = lds128(a00), t1 = lds128(a00 + 128u), t2 = lds128(a00 + WW * 128u),
                      t3 = lds128(a00 + WW * 128u + 128u);
          v[b][0][0] = t0.x; v[b][0][1] = t0.y; v[b][0][2] = t0.z; v[b][0][3] = t0.w;
          v[b][1][0] = t1.x; v[b][1][1] = t1.y; v[b][1][2] = t1.z; v[b][1][3] = t1.w;
          v[b][2][0] = t2.x; v[b][2][1] = t2.y; v[b][2][2] = t2.z; v[b][2][3] = t2.w;
          v[b][3][0] = t3.x; v[b][3][1] = t3.y; v[b][3][2] = t3.z; v[b][3][3] = t3.w;
          }
        } else {                                               // far sample: global gather with explicit validity
          wy0f = ((unsigned)y0 < (unsigned)H) ? wy0f : 0.f;
          wy1f = ((unsigned)y0 + 1u < (unsigned)H) ? wy1f : 0.f;
          wx0f = ((unsigned)x0 < (unsigned)W) ? wx0f : 0.f;
          wx1f = ((unsigned)x0 + 1u < (unsigned)W) ? wx1f : 0.f;
          const int ys = min(max(y0, -1), H), xs = min(max(x0, -1), W);      // keep the +1 below defined
          const int cy0 = min(max(ys, 0), H - 1), cy1 = min(max(ys + 1, 0), H - 1);
          const int cx0 = min(max(xs, 0), W - 1), cx1 = min(max(xs + 1, 0), W - 1);
          const uint32_t b00 = (uint32_t)(cy0 * W + cx0) * CH + (uint32_t)((l ^ u) * 8);
          const uint32_t sx = (uint32_t)(cx1 - cx0) * CH, sy = (uint32_t)((cy1 - cy0) * W) * CH;
          const uint32_t bo[4] = {b00, (uint32_t)(b00 + sx), (uint32_t)(b00 + sy), (uint32_t)(b00 + sy + sx)};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 t = __ldg(reinterpret_cast<const uint4*>(xn + bo[c]));
            v[b][c][0] = t.x; v[b][c][1] = t.y; v[b][c][2] = t.z; v[b][c][3] = t.w;
          }
        }
        const float w00 = wy0f * wx0f, w01 = wy0f * wx1f, w10 = wy1f * wx0f, w11 = wy1f * wx1f;
        if (BLEND16) {
          wq[b][0] = pack_bf16x2(w00, w00); wq[b][1] = pack_bf16x2(w01, w01);
          wq[b][2] = pack_bf16x2(w10, w10); wq[b][3] = pack_bf16x2(w11, w11);
        } else {
          wq[b][0] = __float_as_uint(w00); wq[b][1] = __float_as_uint(w01);
          wq[b][2] = __float_as_uint(w10); wq[b][3] = __float_as_uint(w11);
        }
      };
      auto blend = [&](const int u) {
        const int b = u & 1;
        if (EAVSR_ABL3 & 8) {
#pragma unroll
          for (int e = 0; e < 4; ++e) res[u][e] = v[b][0][e] ^ v[b][1][e] ^ v[b][2][e] ^ v[b][3][e] ^ wq[b][e];
        } else if (BLEND16) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            res[u][e] = hfma2_bf16(wq[b][3], v[b][3][e],
                                   hfma2_bf16(wq[b][2], v[b][2][e], hfma2_bf16(wq[b][1], v[b][1][e], hmul2_bf16(wq[b][0], v[b][0][e]))));
        } else {
          const float w00 = __uint_as_float(wq[b][0]), w01 = __uint_as_float(wq[b][1]);
          const float w10 = __uint_as_float(wq[b][2]), w11 = __uint_as_float(wq[b][3]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lo = w00 * bf16lo_to_f32(v[b][0][e]) + w01 * bf16lo_to_f32(v[b][1][e]) +
                             w10 * bf16lo_to_f32(v[b][2][e]) + w11 * bf16lo_to_f32(v[b][3][e]);
            const float hi = w00 * bf16hi_to_f32(v[b][0][e]) + w01 * bf16hi_to_f32(v[b][1][e]) +
                             w10 * bf16hi_to_f32(v[b][2][e]) + w11 * bf16hi_to_f32(v[b][3][e]);
            res[u][e] = pack_bf16x2(lo, hi);
          }
        }
      };
      issue(0);
#pragma unroll
      for (int u = 0; u < DG; ++u) {
        if (u + 1 < DG) issue(u + 1);
        blend(u);
      }
      // position u holds group l ^ u.  Bit 0 of the permutation is undone here (one stage of an XOR butterfly of
      // selects); bits 1 and 2 move whole 16-channel MMA K-blocks and are undone by the MMA issuer's row classes.
#pragma unroll
      for (int bit = (EAVSR_ABL3 & 128) ? DG : 1; bit < 2; bit <<= 1) {
        const bool sw = (l & bit) != 0;
#pragma unroll
        for (int u = 0; u < DG; ++u) {
          if (u & bit) continue;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t a = res[u][e], c = res[u ^ bit][e];
            res[u][e] = sw ? c : a;
            res[u ^ bit][e] = sw ? a : c;
          }
        }
      }
      if (it >= NSA) mbar_wait_park(bar_aempty + 8 * as, aph ^ 1);  // the MMAs that read this ring slot have completed
      tc_fence_after();
      tmem_st_32x32(tmem_d + ((uint32_t)(qd * 32) << 16) + A_COL0 + as * A_COLS, &res[0][0]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      const bool last_in_tile = tap + SETS >= TAPS;
      if (lane == 0) {
        mbar_arrive(bar_afull + 8 * as);
        if (last_in_tile) mbar_arrive(bar_wine + 8 * (tl & 1));          // this warp is done with the window
      }
      // epilogue of the previous tile, two taps into this one: its last MMAs have long completed by then
      if (++taps_in_tile == 2 && tl >= 1) { epilogue(tl - 1); ep_done = tl; }
      tap += SETS;
      if (tap >= TAPS) { tap -= TAPS; ++tl; newtile = true; }
      os += SETS;
      if (os >= NOS) { os -= NOS; oph ^= 1; }
      as += SETS;
      if (as >= NSA) { as -= NSA; aph ^= 1; }
    }
    for (int t = ep_done; t < my_tiles; ++t) epilogue(t);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PWARPS) tmem_dealloc<TMEM_COLS>(tmem_d);
}

}  // namespace win3
}  // namespace eavsr
