// Warp-specialised tcgen05 DCNv2 forward (second generation of dcn_fwd_tc_kernel), included by
// dcn_fwd_tc.cu.  Same math and data layout; what changed is the schedule:
//
//   * the first kernel put two __syncthreads() around every tap, so all 8 warps drained their loads
//     together and restarted from an empty memory pipeline (ncu: 24 % warps active, 15 % DRAM, 50 %
//     L1, 61 % issue -- bound by none of them, i.e. latency);
//   * here 8 producer warps free-run: each owns 16 rows of the 128-pixel tile, stages its own
//     offsets/masks with cp.async (warp-private, double buffered), keeps the corner loads of the next
//     half-batch in flight while it blends the current one, and hands a finished A stage to the MMA
//     warp through mbarriers (full/empty ring of 3 stages).  No CTA-wide barrier inside the loop;
//   * a ninth warp's elected lane streams the per-tap packed weight tile with cp.async.bulk (TMA
//     engine, mbarrier complete_tx), issues the tcgen05.mma and commits;
//   * two TMEM accumulators (2 x 64 columns): the producers read tile i back (tcgen05.ld) and store
//     it while the MMA warp is already accumulating tile i+1.
#pragma once
#include "common.cuh"

namespace eavsr {
namespace ws {

constexpr int PRODUCER_WARPS = 8;
constexpr int THREADS = (PRODUCER_WARPS + 1) * 32;
constexpr int NS = 3;               // A / B ring depth
constexpr int TILE_M = 128;
constexpr int CH = 64;
constexpr int TAPS = 9;
constexpr int A_TILE_BYTES = TILE_M * CH * 2;
constexpr int B_TILE_BYTES = CH * CH * 2;
constexpr int PLW = 20;             // words per staged offset plane (16 rows + 4 pad: conflict-free)
constexpr int MAX_PLANES = 24;      // 3 * DG, DG <= 8
constexpr int OFF_WARP_BUF = MAX_PLANES * PLW * 4;
constexpr int TMEM_COLS = 128;

template <bool SPLIT> struct Smem {
  static constexpr int TERMS = SPLIT ? 2 : 1;
  static constexpr int A_STAGE = A_TILE_BYTES * TERMS;
  static constexpr int B_STAGE = B_TILE_BYTES * TERMS;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = A_OFF + NS * A_STAGE;
  static constexpr int OFFS_OFF = B_OFF + NS * B_STAGE;
  static constexpr int BAR_OFF = OFFS_OFF + PRODUCER_WARPS * 2 * OFF_WARP_BUF;
  // barriers (8 B each): full[NS], empty[NS], bfull[NS], acc_full[2], acc_empty[2]; then the TMEM slot
  static constexpr int TOTAL = BAR_OFF + (3 * NS + 4) * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

template <typename XT, bool SPLIT, int DG, bool VEC_OFF>
__global__ void __launch_bounds__(THREADS, SPLIT ? 1 : 2)
dcn_fwd_ws_kernel(const XT* __restrict__ x, const float* __restrict__ offset, const float* __restrict__ mask,
                  const uint8_t* __restrict__ wpacked, const XT* __restrict__ bias, XT* __restrict__ out, int H,
                  int W, long long xs_n, long long os_n, int tiles_per_img, int total_tiles) {
  static_assert(DG <= 8, "warp-specialised kernel: deform_groups <= 8");
  using SM = Smem<SPLIT>;
  constexpr int CPI = 8;                                   // channels per item (16 B of bf16)
  constexpr int NPLANES = 3 * DG;
  constexpr int HB = 1;                                    // items per batch (2 batches in flight)
  constexpr int NB = 4 / HB;                               // batches per tap
  constexpr int NW = RawVec<XT, CPI>::NW;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base + SM::A_OFF, sB = smem_base + SM::B_OFF;
  const uint32_t bars = smem_base + SM::BAR_OFF;
  const uint32_t bar_full = bars, bar_empty = bars + NS * 8, bar_bfull = bars + 2 * NS * 8;
  const uint32_t bar_accf = bars + 3 * NS * 8, bar_acce = bar_accf + 16;
  const uint32_t tmem_slot_addr = bar_acce + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM::BAR_OFF + (3 * NS + 4) * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = H * W;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bar_full + 8 * s, PRODUCER_WARPS);
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_bfull + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accf + 8 * b, 1);
      mbar_init(bar_acce + 8 * b, PRODUCER_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == PRODUCER_WARPS) tmem_alloc<TMEM_COLS>(tmem_slot_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  const int first_tile = blockIdx.x;
  const int my_tiles = (total_tiles - first_tile + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_iters = my_tiles * TAPS;
  constexpr uint32_t IDESC = umma_idesc_bf16(TILE_M, CH);

  if (warp == PRODUCER_WARPS) {
    // ================= MMA issuer + weight-tile loader (one elected lane) =================
    if (elect_one()) {
      auto issue_b = [&](int j) {
        const uint32_t bar = bar_bfull + 8 * (j % NS);
        mbar_arrive_expect_tx(bar, SM::B_STAGE);
        bulk_g2s(sB + (j % NS) * SM::B_STAGE, wpacked + (size_t)(j % TAPS) * SM::B_STAGE, SM::B_STAGE, bar);
      };
      for (int j = 0; j < NS && j < n_iters; ++j) issue_b(j);
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % NS, u = it / NS;
        const int tap = it % TAPS, tl = it / TAPS, buf = tl & 1;
        if (it >= 1 && it - 1 + NS < n_iters) {            // stage of it-1 is free once MMA(it-1) retired
          mbar_wait(bar_empty + 8 * ((it - 1) % NS), ((it - 1) / NS) & 1);
          issue_b(it - 1 + NS);
        }
        if (tap == 0 && tl >= 2) mbar_wait(bar_acce + 8 * buf, ((tl >> 1) - 1) & 1);
        mbar_wait(bar_full + 8 * s, u & 1);
        mbar_wait(bar_bfull + 8 * s, u & 1);
        tc_fence_after();
        const uint32_t aStage = sA + s * SM::A_STAGE, bStage = sB + s * SM::B_STAGE;
        const uint64_t a_hi = umma_desc_sw128_kmajor(aStage), b_hi = umma_desc_sw128_kmajor(bStage);
        const uint32_t d = tmem_d + buf * CH;
#pragma unroll
        for (int k = 0; k < CH / 16; ++k) {
          umma_bf16(d, a_hi + 2 * k, b_hi + 2 * k, IDESC, (tap | k) != 0);
          if (SPLIT) {
            const uint64_t a_lo = umma_desc_sw128_kmajor(aStage + A_TILE_BYTES);
            const uint64_t b_lo = umma_desc_sw128_kmajor(bStage + B_TILE_BYTES);
            umma_bf16(d, a_lo + 2 * k, b_hi + 2 * k, IDESC, 1);
            umma_bf16(d, a_hi + 2 * k, b_lo + 2 * k, IDESC, 1);
          }
        }
        umma_commit(bar_empty + 8 * s);
        if (tap == TAPS - 1) umma_commit(bar_accf + 8 * buf);
      }
    }
    __syncwarp();
  } else {
    // ================= producer warps: gather -> blend -> swizzled A stage =================
    const int q = lane >> 3, l = lane & 7;                  // row within a group of 4, channel chunk
    const int grp = (l * DG) / 8;
    const uint32_t offBase = smem_base + SM::OFFS_OFF + warp * 2 * OFF_WARP_BUF;
    const float* offF = reinterpret_cast<const float*>(smem + SM::OFFS_OFF + warp * 2 * OFF_WARP_BUF);

    auto tile_of = [&](int it) { return first_tile + (it / TAPS) * (int)gridDim.x; };

    auto prefetch_offsets = [&](int it) {                   // stage offsets/mask rows of this warp
      if (it < n_iters) {
        const int tile = tile_of(it), tap = it % TAPS;
        const int n = tile / tiles_per_img;
        const int pix0 = (tile - n * tiles_per_img) * TILE_M + 4 * warp;
        const uint32_t dst0 = offBase + (it & 1) * OFF_WARP_BUF;
        constexpr int PER_PLANE = VEC_OFF ? 4 : 16;
        for (int i = lane; i < NPLANES * PER_PLANE; i += 32) {
          const int plane = i / PER_PLANE, e = i - plane * PER_PLANE;
          const int comp = plane / DG, g = plane - comp * DG;
          const int lr = VEC_OFF ? e * 4 : e;               // local row (j*4 + q')
          const int pix = pix0 + (lr >> 2) * 32 + (lr & 3);
          if (pix < HW) {
            const float* src = (comp < 2)
                ? offset + ((size_t)(n * DG + g) * TAPS + tap) * 2 * HW + (size_t)comp * HW + pix
                : mask + ((size_t)(n * DG + g) * TAPS + tap) * HW + pix;
            const uint32_t dst = dst0 + (plane * PLW + lr) * 4;
            if (VEC_OFF) cp_async_16(dst, src); else cp_async_4(dst, src);
          }
        }
      }
      cp_async_commit();                                    // always commit: uniform group counting
    };

    uint32_t raw[2][HB][4][NW];
    float wgt[2][HB][4];
    float py_f[4], px_f[4];
    const XT* xn = x;
    int cur_tile = -1;

    // compute addresses + weights of batch b of iteration `it` and put its loads in flight
    auto issue = [&](int it, int b, int pb) {
      const int tile = tile_of(it), tap = it % TAPS;
      if (tile != cur_tile) {
        cur_tile = tile;
        const int n = tile / tiles_per_img;
        const int pix0 = (tile - n * tiles_per_img) * TILE_M;
        xn = x + (size_t)n * xs_n;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int pix = pix0 + 4 * warp + q + 32 * j;
          const int yy = pix / W;
          py_f[j] = pix < HW ? (float)(yy - 1) : -100000.f;
          px_f[j] = pix < HW ? (float)(pix - yy * W - 1) : 0.f;
        }
      }
      const float* so = offF + (it & 1) * (OFF_WARP_BUF / 4);
      const int ti = tap / 3, tj = tap - ti * 3;
#pragma unroll
      for (int jj = 0; jj < HB; ++jj) {
        const int j = b * HB + jj;
        const int lr = j * 4 + q;
        const float dy = so[(0 * DG + grp) * PLW + lr];
        const float dx = so[(1 * DG + grp) * PLW + lr];
        const float m = so[(2 * DG + grp) * PLW + lr];
        const float py = (py_f[j] + (float)ti) + dy;
        const float px = (px_f[j] + (float)tj) + dx;
        const float fy = floorf(py), fx = floorf(px);
        const float ly = py - fy, lx = px - fx;
        // cell index, saturated so that wild / NaN offsets land on "all corners outside"
        const int y0 = (int)fminf(fmaxf(fy, -2.f), (float)H), x0 = (int)fminf(fmaxf(fx, -2.f), (float)W);
        // separable weights with the mask folded into the row weights.  A row / column outside the
        // image gets weight 0, which also implements DCNv2's "p <= -1 or p >= size -> 0" rule
        // (exactly on p == -1 the only in-image corner has weight l == 0).
        const float wy0 = ((unsigned)y0 < (unsigned)H) ? m * (1.f - ly) : 0.f;
        const float wy1 = ((unsigned)(y0 + 1) < (unsigned)H) ? m * ly : 0.f;
        const float wx0 = ((unsigned)x0 < (unsigned)W) ? (1.f - lx) : 0.f;
        const float wx1 = ((unsigned)(x0 + 1) < (unsigned)W) ? lx : 0.f;
        wgt[pb][jj][0] = wy0 * wx0;
        wgt[pb][jj][1] = wy0 * wx1;
        wgt[pb][jj][2] = wy1 * wx0;
        wgt[pb][jj][3] = wy1 * wx1;
        const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
        const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1);
        const uint32_t b00 = (uint32_t)(cy0 * W + cx0) * CH + l * CPI;
        const uint32_t sxo = (uint32_t)(cx1 - cx0) * CH, syo = (uint32_t)((cy1 - cy0) * W) * CH;
        RawVec<XT, CPI>::ld(xn + b00, raw[pb][jj][0]);
        RawVec<XT, CPI>::ld(xn + (uint32_t)(b00 + sxo), raw[pb][jj][1]);
        RawVec<XT, CPI>::ld(xn + (uint32_t)(b00 + syo), raw[pb][jj][2]);
        RawVec<XT, CPI>::ld(xn + (uint32_t)(b00 + syo + sxo), raw[pb][jj][3]);
      }
    };

    auto blend_store = [&](int it, int b, int pb) {
      const uint32_t aStage = sA + (it % NS) * SM::A_STAGE;
#pragma unroll
      for (int jj = 0; jj < HB; ++jj) {
        const int r = 4 * warp + q + 32 * (b * HB + jj);
        float v[CPI];
#pragma unroll
        for (int e = 0; e < CPI; ++e)
          v[e] = wgt[pb][jj][0] * RawVec<XT, CPI>::get(raw[pb][jj][0], e) +
                 wgt[pb][jj][1] * RawVec<XT, CPI>::get(raw[pb][jj][1], e) +
                 wgt[pb][jj][2] * RawVec<XT, CPI>::get(raw[pb][jj][2], e) +
                 wgt[pb][jj][3] * RawVec<XT, CPI>::get(raw[pb][jj][3], e);
        uint32_t hi[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) hi[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
        const uint32_t dst = aStage + sw128_offset(r, l * 16);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]));
        if (SPLIT) {
          uint32_t lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            lo[e] = pack_bf16x2(v[2 * e] - bf16lo_to_f32(hi[e]), v[2 * e + 1] - bf16hi_to_f32(hi[e]));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst + A_TILE_BYTES), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]));
        }
      }
    };

    auto epilogue = [&](int tl) {                           // TMEM accumulator of local tile tl -> out
      const int buf = tl & 1;
      mbar_wait(bar_accf + 8 * buf, (tl >> 1) & 1);
      tc_fence_after();
      const int qd = warp & 3, half = warp >> 2;
      uint32_t acc[32];
      tmem_ld_32x32(tmem_d + ((uint32_t)(qd * 32) << 16) + buf * CH + half * 32, acc);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce + 8 * buf);
      const int tile = first_tile + tl * (int)gridDim.x;
      const int n = tile / tiles_per_img;
      const int pix = (tile - n * tiles_per_img) * TILE_M + qd * 32 + lane;
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
    };

    prefetch_offsets(0);
    prefetch_offsets(1);
    cp_async_wait<1>();
    __syncwarp();
    issue(0, 0, 0);
    for (int it = 0; it < n_iters; ++it) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        if (b + 1 < NB) {
          issue(it, b + 1, (b + 1) & 1);
          if (b + 2 == NB) {                                // all addresses of `it` are computed
            __syncwarp();
            prefetch_offsets(it + 2);                       // reuses offset buffer (it & 1)
          }
        } else {
          cp_async_wait<1>();                               // offsets of it+1 have landed
          __syncwarp();
          if (it + 1 < n_iters) issue(it + 1, 0, 0);
        }
        if (b == 0 && it >= NS) mbar_wait(bar_empty + 8 * (it % NS), ((it / NS) - 1) & 1);
        blend_store(it, b, b & 1);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * (it % NS));
      if (it % TAPS == 1 && it >= TAPS) epilogue(it / TAPS - 1);
    }
    cp_async_wait<0>();
    epilogue(my_tiles - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PRODUCER_WARPS) tmem_dealloc<TMEM_COLS>(tmem_d);
}

}  // namespace ws
}  // namespace eavsr
