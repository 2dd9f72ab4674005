// DCNv2 backward on tcgen05 / TMEM (sm_100a) for the model's configuration: 64 -> 64 channels, 3x3,
// stride 1, pad 1, dilation 1, groups 1, deform_groups in {1,2,4,8,16}, bf16 NHWC features and gradients.
//
// Replaces, for that configuration, what mmcv runs under the autograd of models/networks.py:627-630:
// modulated_deformable_col2im + col2im_coord over a 576 x P fp32 column-gradient buffer in HBM and
// two cuBLAS GEMMs.  Here neither the column buffer nor its gradient ever exists in HBM:
//
//   data kernel   dcol_tap[128 px x 64 cin] = gout[128 px x 64 cout] * W_tap^T        (tcgen05, TMEM)
//                 each thread owns one pixel (its TMEM lane) and 16 input channels: it re-samples x from
//                 a shared-memory window and turns dcol into d(mask), d(offset) and the bilinear scatter
//                 of d(x) (fp32 vector reductions to global memory) -- no col2im pass.
//   weight kernel dW_tap[64 cin x 64 cout] += col_tap^T[64 x 128 px] * gout[128 px x 64 cout]
//                 col_tap is produced exactly like in the forward kernel (window gather -> bf16 swizzled
//                 tile); both operands are consumed MN-major (the pixel dimension is the contraction),
//                 so no transposed copy is needed.  The nine 64x64 fp32 accumulators of a CTA stay in
//                 TMEM over all of its tiles (two M=64 tiles share 64 columns: lanes 0-15 / 16-31 of every
//                 sub-partition) and are reduced into gweight once at the end.
//
// Descriptors: cute::UMMA::SmemDescriptor / InstrDescriptor bit layouts (see common.cuh); an MN-major
// SWIZZLE_128B operand is ((8 x 16 B contiguous in MN), (8 rows of 128 B, SBO = 1024 B between row groups)).
#include "common.cuh"
#include "dcn_fwd_win.cuh"

namespace eavsr {
namespace bwd {

using win::lds128;
using win::lds64;

constexpr int CWARPS = 16;
constexpr int THREADS = (CWARPS + 1) * 32;      // 16 worker warps + 1 MMA / loader warp
constexpr int TH = 8, TW = 16;                  // tile = 128 pixels = UMMA M (data) / K (weight)
constexpr int PAD = 5;
constexpr int WH = TH + 2 * PAD, WW = TW + 2 * PAD;
constexpr int WIN_BYTES = WH * WW * 128;        // 59 904
constexpr int CH = 64, TAPS = 9;
constexpr int G_TILE = 128 * CH * 2;            // gout tile, 16 KB
constexpr int B_TILE = CH * CH * 2;             // W_tap^T, 8 KB

__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// MN-major SWIZZLE_128B operand whose MN extent is one 64-element atom (LBO unused)
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t IDESC_MN_MAJOR = (1u << 15) | (1u << 16);

// Packed transposed weights for the data kernel: [tap][8 KB K-major SW128 tile], row = cin, k = cout.
__global__ void dcn_pack_weight_t(const __nv_bfloat16* __restrict__ w, uint8_t* __restrict__ packed) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (tap, c, o)
  if (idx >= TAPS * CH * CH) return;
  const int o = idx % CH, c = (idx / CH) % CH, t = idx / (CH * CH);
  *reinterpret_cast<__nv_bfloat16*>(packed + (size_t)t * B_TILE + sw128_offset(c, o * 2)) = w[((size_t)o * CH + c) * TAPS + t];
}

// d(bias)[o] = sum over pixels of gout[., o]: gout is NHWC, so a thread owns one 16-byte chunk (8 channels)
// and strides over pixels; 256 threads = 32 pixels x 8 chunks per step, shared-memory tree, 64 atomics per CTA.
__global__ void __launch_bounds__(256)
dcn_bwd_bias_nhwc(const __nv_bfloat16* __restrict__ gout, float* __restrict__ gbias, int n, long long HW,
                  long long gs_n) {
  __shared__ float red[32][CH + 1];
  const int c = threadIdx.x & 7, p = threadIdx.x >> 3;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long total = (long long)n * HW;
  for (long long i = (long long)blockIdx.x * 32 + p; i < total; i += (long long)gridDim.x * 32) {
    const long long img = i / HW, pix = i - img * HW;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(gout + img * gs_n + pix * CH + c * 8));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { a[2 * e] += bf16lo_to_f32(w[e]); a[2 * e + 1] += bf16hi_to_f32(w[e]); }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[p][c * 8 + e] = a[e];
  __syncthreads();
  if (threadIdx.x < CH) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    atomicAdd(gbias + threadIdx.x, s);
  }
}

// ==================================================================================================
// data kernel: d(x), d(offset), d(mask)
// ==================================================================================================
namespace data {
constexpr int NSB = 2, NACC = 4;
constexpr int TMEM_COLS = NACC * CH;            // 256
struct Smem {
  static constexpr int WIN_OFF = 0;
  static constexpr int G_OFF = WIN_OFF + 2 * WIN_BYTES;
  static constexpr int B_OFF = G_OFF + 2 * G_TILE;
  static constexpr int BAR_OFF = B_OFF + NSB * B_TILE;
  static constexpr int NBARS = 2 + NSB + 2 * NACC;   // gfull[2], bfull[NSB], dfull[NACC], dempty[NACC]
  static constexpr int TOTAL = BAR_OFF + NBARS * 8 + 16;
  static constexpr int DYN = TOTAL + 1024;
};
static_assert(Smem::G_OFF % 1024 == 0 && Smem::B_OFF % 1024 == 0, "operand tiles must be 1024-byte aligned");

template <int DG>
__global__ void __launch_bounds__(THREADS, 1)
dcn_bwd_data_kernel(const __nv_bfloat16* __restrict__ gout, const __nv_bfloat16* __restrict__ x,
                    const float* __restrict__ offset, const float* __restrict__ mask,
                    const uint8_t* __restrict__ wtpacked, float* __restrict__ gx32, float* __restrict__ goffset,
                    float* __restrict__ gmask, int H, int W, long long xs_n, long long gs_n, long long gxs_n,
                    int tiles_x, int tiles_per_img, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t sWin = sbase + Smem::WIN_OFF, sG = sbase + Smem::G_OFF, sB = sbase + Smem::B_OFF;
  const uint32_t bars = sbase + Smem::BAR_OFF;
  const uint32_t bar_gfull = bars, bar_bfull = bar_gfull + 16, bar_dfull = bar_bfull + NSB * 8;
  const uint32_t bar_dempty = bar_dfull + NACC * 8;
  const uint32_t tmem_slot_addr = bar_dempty + NACC * 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Smem::BAR_OFF + Smem::NBARS * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = H * W;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) mbar_init(bar_gfull + 8 * b, CWARPS);
    for (int s = 0; s < NSB; ++s) mbar_init(bar_bfull + 8 * s, 1);
    for (int a = 0; a < NACC; ++a) { mbar_init(bar_dfull + 8 * a, 1); mbar_init(bar_dempty + 8 * a, CWARPS); }
    fence_mbar_init();
  }
  if (warp == CWARPS) tmem_alloc<TMEM_COLS>(tmem_slot_addr);
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

  if (warp == CWARPS) {
    // ============ MMA issuer + weight-tile loader (one lane) ============
    if (elect_one()) {
      auto issue_b = [&](int j) {
        const uint32_t bar = bar_bfull + 8 * (j % NSB);
        mbar_arrive_expect_tx(bar, B_TILE);
        bulk_g2s(sB + (j % NSB) * B_TILE, wtpacked + (size_t)(j % TAPS) * B_TILE, B_TILE, bar);
      };
      for (int j = 0; j < NSB && j < n_iters; ++j) issue_b(j);
      const uint64_t g_base = umma_desc_sw128_kmajor(sG), b_base = umma_desc_sw128_kmajor(sB);
      int tap = 0, tl = 0;
      for (int it = 0; it < n_iters; ++it) {
        const int acc = it % NACC, sb = it % NSB, buf = tl & 1;
        if (tap == 0) mbar_wait(bar_gfull + 8 * buf, (tl >> 1) & 1);
        mbar_wait(bar_bfull + 8 * sb, (it / NSB) & 1);
        if (it >= NACC) mbar_wait(bar_dempty + 8 * acc, ((it / NACC) - 1) & 1);
        tc_fence_after();
        const uint64_t a_d = g_base + (uint64_t)((buf * G_TILE) >> 4), b_d = b_base + (uint64_t)((sb * B_TILE) >> 4);
        const uint32_t d = tmem_d + acc * CH;
#pragma unroll
        for (int k = 0; k < CH / 16; ++k) umma_bf16(d, a_d + 2 * k, b_d + 2 * k, IDESC, k != 0);
        umma_commit(bar_dfull + 8 * acc);
        // refill the weight ring behind the MMAs (stage of it-1 is free once MMA(it-1) retired)
        if (it >= 1 && it - 1 + NSB < n_iters) {
          mbar_wait(bar_dfull + 8 * ((it - 1) % NACC), ((it - 1) / NACC) & 1);
          issue_b(it - 1 + NSB);
        }
        if (++tap == TAPS) { tap = 0; ++tl; }
      }
    }
    __syncwarp();
  } else {
    // ============ workers: dcol (TMEM) -> d(mask), d(offset), scatter of d(x) ============
    const int qd = warp & 3, cq = warp >> 2;                  // TMEM lane quadrant, 16-channel quarter
    const int m = qd * 32 + lane;                             // pixel of the tile == TMEM lane
    const int prow = m >> 4, pcol = m & 15;

    // window (swizzled: 16-byte chunk c of cell i sits at chunk c ^ (i & 7), so that the 32 lanes of a
    // warp -- 32 neighbouring pixels reading the same chunk -- hit different banks) and gout tile of a tile
    auto fill = [&](int tl) {
      int n, ty0, tx0;
      tile_coords(tl, n, ty0, tx0);
      const int wy0 = ty0 - PAD, wx0 = tx0 - PAD;
      const __nv_bfloat16* xn = x + (size_t)n * xs_n;
      const uint32_t wdst = sWin + (tl & 1) * WIN_BYTES;
      for (int i = tid; i < WH * WW * 8; i += CWARPS * 32) {
        const int cell = i >> 3, c = i & 7;
        const int cy = wy0 + cell / WW, cx = wx0 + cell % WW;
        const bool ok = (unsigned)cy < (unsigned)H && (unsigned)cx < (unsigned)W;
        const __nv_bfloat16* src = ok ? xn + ((size_t)cy * W + cx) * CH + c * 8 : xn;
        cp_async_16_zfill(wdst + cell * 128 + ((c ^ (cell & 7)) << 4), src, ok);
      }
      const __nv_bfloat16* gn = gout + (size_t)n * gs_n;
      const uint32_t gdst = sG + (tl & 1) * G_TILE;
      for (int i = tid; i < 128 * 8; i += CWARPS * 32) {
        const int r = i >> 3, c = i & 7;
        const int gy = ty0 + (r >> 4), gx = tx0 + (r & 15);
        const bool ok = gy < H && gx < W;
        const __nv_bfloat16* src = ok ? gn + ((size_t)gy * W + gx) * CH + c * 8 : gn;
        cp_async_16_zfill(gdst + sw128_offset(r, c * 16), src, ok);
      }
    };

    // offsets / mask of this thread's pixel for the groups of its two 8-channel chunks
    constexpr int NSUB = DG == 16 ? 2 : 1;    // groups (samples) per chunk
    constexpr int NS = 2 * NSUB;              // samples per thread and tap
    constexpr int CPS = 8 / NSUB;             // channels per sample
    constexpr int CW = CPS / 2;               // 32-bit words per corner
    struct Om { float dy[NS], dx[NS], mk[NS]; };
    auto load_om = [&](int tl, int tap) {
      Om o;
      int n, ty0, tx0;
      tile_coords(tl, n, ty0, tx0);
      const int gy = ty0 + prow, gx = tx0 + pcol;
      const bool live = gy < H && gx < W;
#pragma unroll
      for (int u = 0; u < NS; ++u) {
        const int jc = 2 * cq + u / NSUB;
        const int g = NSUB == 2 ? 2 * jc + u % NSUB : (jc * DG) >> 3;
        const size_t pix = (size_t)gy * W + gx;
        const size_t ob = ((size_t)(n * DG + g) * TAPS + tap) * 2 * HW + pix;
        const size_t mb = ((size_t)(n * DG + g) * TAPS + tap) * HW + pix;
        o.dy[u] = live ? __ldg(offset + ob) : 0.f;
        o.dx[u] = live ? __ldg(offset + ob + HW) : 0.f;
        o.mk[u] = live ? __ldg(mask + mb) : 0.f;
      }
      return o;
    };

    fill(0);
    cp_async_commit();
    Om cur = load_om(0, 0);
    int it = 0;
    for (int tl = 0; tl < my_tiles; ++tl) {
      int n, ty0, tx0;
      tile_coords(tl, n, ty0, tx0);
      cp_async_wait<0>();
      if (tl == 0) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_gfull);
      }
      // window of this tile visible to everyone; everyone has left tile tl-1 (its window / gout buffers
      // may be refilled: all MMAs that read that gout tile have completed, the workers consumed them)
      asm volatile("bar.sync 1, %0;\n" ::"n"(CWARPS * 32) : "memory");
      if (tl + 1 < my_tiles) fill(tl + 1);
      cp_async_commit();

      const __nv_bfloat16* xn = x + (size_t)n * xs_n;
      const int wy0 = ty0 - PAD, wx0 = tx0 - PAD;
      const uint32_t win = sWin + (tl & 1) * WIN_BYTES;
      const int gy = ty0 + prow, gx = tx0 + pcol;
      const bool live = gy < H && gx < W;
      const size_t pix = (size_t)gy * W + gx;
      float* gxn = gx32 ? gx32 + (size_t)n * gxs_n : nullptr;

      for (int tap = 0; tap < TAPS; ++tap, ++it) {
        if (tap == 4 && tl + 1 < my_tiles) {                   // gout tile of the next tile has landed
          cp_async_wait<0>();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_gfull + 8 * ((tl + 1) & 1));
        }
        Om nxt;
        {
          const int ntap = tap + 1 == TAPS ? 0 : tap + 1, ntl = tap + 1 == TAPS ? tl + 1 : tl;
          if (ntl < my_tiles) nxt = load_om(ntl, ntap); else nxt = cur;
        }
        const int acc = it % NACC;
        mbar_wait(bar_dfull + 8 * acc, (it / NACC) & 1);
        tc_fence_after();
        uint32_t dr[16];
        tmem_ld_x16(tmem_d + ((uint32_t)(qd * 32) << 16) + acc * CH + cq * 16, dr);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dempty + 8 * acc);

        if (live) {
          const int ti = (tap * 11) >> 5, tj = tap - ti * 3;
#pragma unroll
          for (int u = 0; u < NS; ++u) {
            const int j = u / NSUB, sb = u % NSUB;
            const int jc = 2 * cq + j;                         // 8-channel chunk
            const int g = NSUB == 2 ? 2 * jc + sb : (jc * DG) >> 3;
            const int d0 = 8 * j + sb * CPS;                   // first dcol register of this sample
            const float mk = cur.mk[u];
            const float py = (float)(gy - 1 + ti) + cur.dy[u], px = (float)(gx - 1 + tj) + cur.dx[u];
            const bool inside = py > -1.f && py < (float)H && px > -1.f && px < (float)W;   // mmcv's rule
            float dm = 0.f, gyv = 0.f, gxv = 0.f;
            if (inside) {
              const int y0 = __float2int_rd(py), x0 = __float2int_rd(px);
              const float ly = py - (float)y0, lx = px - (float)x0;
              const int ry = y0 - wy0, rx = x0 - wx0;
              const bool vy0 = y0 >= 0, vy1 = y0 + 1 < H, vx0 = x0 >= 0, vx1 = x0 + 1 < W;
              uint32_t v[4][CW];
              if ((unsigned)ry < (unsigned)(WH - 1) && (unsigned)rx < (unsigned)(WW - 1)) {
                const int c00 = ry * WW + rx;
                const int cells[4] = {c00, c00 + 1, c00 + WW, c00 + WW + 1};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint32_t ad = win + cells[q] * 128 + ((jc ^ (cells[q] & 7)) << 4) + sb * 8;
                  if (CW == 4) {
                    const uint4 t = lds128(ad);
                    v[q][0] = t.x; v[q][1] = t.y; v[q][CW - 2] = t.z; v[q][CW - 1] = t.w;
                  } else {
                    const uint2 t = lds64(ad);
                    v[q][0] = t.x; v[q][1] = t.y;
                  }
                }
              } else {                                         // far sample: global gather, zero where invalid
                const int cy0 = max(y0, 0), cy1 = min(y0 + 1, H - 1), cx0 = max(x0, 0), cx1 = min(x0 + 1, W - 1);
                const __nv_bfloat16* xb = xn + jc * 8 + sb * CPS;
                const size_t po[4] = {(size_t)cy0 * W + cx0, (size_t)cy0 * W + cx1, (size_t)cy1 * W + cx0, (size_t)cy1 * W + cx1};
                const bool okq[4] = {vy0 && vx0, vy0 && vx1, vy1 && vx0, vy1 && vx1};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  if (CW == 4) {
                    const uint4 t = okq[q] ? __ldg(reinterpret_cast<const uint4*>(xb + po[q] * CH)) : make_uint4(0, 0, 0, 0);
                    v[q][0] = t.x; v[q][1] = t.y; v[q][CW - 2] = t.z; v[q][CW - 1] = t.w;
                  } else {
                    const uint2 t = okq[q] ? __ldg(reinterpret_cast<const uint2*>(xb + po[q] * CH)) : make_uint2(0, 0);
                    v[q][0] = t.x; v[q][1] = t.y;
                  }
                }
              }
#pragma unroll
              for (int e = 0; e < CPS; ++e) {
                const float fa = (e & 1) ? bf16hi_to_f32(v[0][e >> 1]) : bf16lo_to_f32(v[0][e >> 1]);
                const float fb = (e & 1) ? bf16hi_to_f32(v[1][e >> 1]) : bf16lo_to_f32(v[1][e >> 1]);
                const float fc = (e & 1) ? bf16hi_to_f32(v[2][e >> 1]) : bf16lo_to_f32(v[2][e >> 1]);
                const float fd = (e & 1) ? bf16hi_to_f32(v[3][e >> 1]) : bf16lo_to_f32(v[3][e >> 1]);
                const float ba = fb - fa, dc = fd - fc;
                const float top = fmaf(lx, ba, fa), bot = fmaf(lx, dc, fc);
                const float bt = bot - top;
                const float val = fmaf(ly, bt, top);
                const float ddx = fmaf(ly, dc - ba, ba);
                const float dcol = __uint_as_float(dr[d0 + e]);
                dm = fmaf(dcol, val, dm);
                gyv = fmaf(dcol, bt, gyv);
                gxv = fmaf(dcol, ddx, gxv);
              }
              if (gxn) {
                const float hy = 1.f - ly, hx = 1.f - lx;
                const float wq[4] = {hy * hx * mk, hy * lx * mk, ly * hx * mk, ly * lx * mk};
                const bool okq[4] = {vy0 && vx0, vy0 && vx1, vy1 && vx0, vy1 && vx1};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  if (okq[q]) {
                    float* dst = gxn + ((size_t)(y0 + (q >> 1)) * W + (x0 + (q & 1))) * CH + jc * 8 + sb * CPS;
                    const float s = wq[q];
#pragma unroll
                    for (int e = 0; e < CPS; e += 4)
                      atomicAdd(reinterpret_cast<float4*>(dst + e),
                                make_float4(s * __uint_as_float(dr[d0 + e]), s * __uint_as_float(dr[d0 + e + 1]),
                                            s * __uint_as_float(dr[d0 + e + 2]), s * __uint_as_float(dr[d0 + e + 3])));
                  }
                }
              }
            }
            const size_t ob = ((size_t)(n * DG + g) * TAPS + tap) * 2 * HW + pix;
            const size_t mb = ((size_t)(n * DG + g) * TAPS + tap) * HW + pix;
            if (DG >= 8) {                                     // one sample per group: plain stores
              if (gmask) gmask[mb] = dm;
              if (goffset) { goffset[ob] = mk * gyv; goffset[ob + HW] = mk * gxv; }
            } else if (inside) {                               // several chunks per group: accumulate
              if (gmask) atomicAdd(gmask + mb, dm);
              if (goffset) { atomicAdd(goffset + ob, mk * gyv); atomicAdd(goffset + ob + HW, mk * gxv); }
            }
          }
        }
        cur = nxt;
      }
    }
    cp_async_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == CWARPS) tmem_dealloc<TMEM_COLS>(tmem_d);
}
}  // namespace data

// ==================================================================================================
// weight kernel: d(weight)
// ==================================================================================================
namespace wgt {
constexpr int NSA = 3, NOB = 2;
constexpr int A_TILE = 128 * CH * 2;
constexpr int PLW = 8;
constexpr int TMEM_COLS = 512;                  // 5 x 64 columns used (two taps per 64 columns)
// deform_groups = 16: two 4-channel groups per chunk, 48 offset planes per tap, 4-pixel apron (as in
// win::Cfg<16>) -- the window is the only thing that can shrink to make room for the planes.
template <int DG> struct Cfg {
  static constexpr int NSUB = DG == 16 ? 2 : 1;
  static constexpr int PAD = DG == 16 ? 4 : 5;
  static constexpr int WH = TH + 2 * PAD, WW = TW + 2 * PAD;
  static constexpr int WIN_BYTES = WH * WW * 128;
  static constexpr int MAX_PLANES = DG == 16 ? 48 : 24;
  static constexpr int OFF_WARP_BUF = MAX_PLANES * PLW * 4;
  static constexpr int WIN_OFF = 0;
  static constexpr int A_OFF = WIN_OFF + 2 * WIN_BYTES;
  static constexpr int G_OFF = A_OFF + NSA * A_TILE;
  static constexpr int OFFS_OFF = G_OFF + 2 * G_TILE;
  static constexpr int BAR_OFF = OFFS_OFF + CWARPS * NOB * OFF_WARP_BUF;
  // full[NSA], empty[NSA], gfull[2], winf[2], wine[2], done
  static constexpr int NBARS = 2 * NSA + 7;
  static constexpr int TOTAL = BAR_OFF + NBARS * 8 + 16;
  static constexpr int DYN = TOTAL + 1024;
  static_assert(A_OFF % 1024 == 0 && G_OFF % 1024 == 0, "operand tiles must be 1024-byte aligned");
  static_assert(DYN <= 232448, "shared memory budget");
};

template <int DG, bool VEC_OFF>
__global__ void __launch_bounds__(THREADS, 1)
dcn_bwd_weight_kernel(const __nv_bfloat16* __restrict__ gout, const __nv_bfloat16* __restrict__ x,
                      const float* __restrict__ offset, const float* __restrict__ mask, float* __restrict__ gweight,
                      int H, int W, long long xs_n, long long gs_n, int tiles_x, int tiles_per_img, int total_tiles) {
  using Smem = Cfg<DG>;
  constexpr int NSUB = Smem::NSUB, PAD = Smem::PAD, WH = Smem::WH, WW = Smem::WW, WIN_BYTES = Smem::WIN_BYTES;
  constexpr int OFF_WARP_BUF = Smem::OFF_WARP_BUF;
  constexpr int NPLANES = 3 * DG;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t sWin = sbase + Smem::WIN_OFF, sA = sbase + Smem::A_OFF, sG = sbase + Smem::G_OFF;
  const uint32_t bars = sbase + Smem::BAR_OFF;
  const uint32_t bar_full = bars, bar_empty = bar_full + NSA * 8, bar_gfull = bar_empty + NSA * 8;
  const uint32_t bar_winf = bar_gfull + 16, bar_wine = bar_winf + 16, bar_done = bar_wine + 16;
  const uint32_t tmem_slot_addr = bar_done + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Smem::BAR_OFF + Smem::NBARS * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = H * W;
  if (tid == 0) {
    for (int s = 0; s < NSA; ++s) { mbar_init(bar_full + 8 * s, CWARPS); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_gfull + 8 * b, CWARPS);
      mbar_init(bar_winf + 8 * b, 1);
      mbar_init(bar_wine + 8 * b, CWARPS);
    }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 2 * WIN_BYTES / 16; i += THREADS)
    *reinterpret_cast<uint4*>(smem + Smem::WIN_OFF + i * 16) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == CWARPS) tmem_alloc<TMEM_COLS>(tmem_slot_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  const int first = blockIdx.x;
  const int my_tiles = (total_tiles - first + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_iters = my_tiles * TAPS;
  // M = 64 (cin), N = 64 (cout), both operands MN-major
  constexpr uint32_t IDESC = umma_idesc_bf16(64, CH) | IDESC_MN_MAJOR;

  auto tile_coords = [&](int tl, int& n, int& ty0, int& tx0) {
    const int tile = first + tl * (int)gridDim.x;
    n = tile / tiles_per_img;
    const int rem = tile - n * tiles_per_img;
    ty0 = (rem / tiles_x) * TH;
    tx0 = (rem % tiles_x) * TW;
  };

  if (warp == CWARPS) {
    // ============ MMA issuer + window loader (one lane) ============
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
      if (my_tiles > 0) load_window(0);
      const uint64_t a_base = umma_desc_sw128_mnmajor(sA), g_base = umma_desc_sw128_mnmajor(sG);
      int tap = 0, tl = 0;
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % NSA, buf = tl & 1;
        mbar_wait(bar_full + 8 * s, (it / NSA) & 1);
        if (tap == 0) {
          mbar_wait(bar_gfull + 8 * buf, (tl >> 1) & 1);
          if (tl + 1 < my_tiles) {                             // workers have left tile tl-1: refill its window
            if (tl >= 1) mbar_wait(bar_wine + 8 * ((tl + 1) & 1), (((tl + 1) >> 1) - 1) & 1);
            load_window(tl + 1);
          }
        }
        tc_fence_after();
        const uint64_t a_d = a_base + (uint64_t)((s * A_TILE) >> 4), b_d = g_base + (uint64_t)((buf * G_TILE) >> 4);
        // accumulator of tap t: columns 64*(t>>1), lanes +16*(t&1) of every sub-partition (M = 64 layout)
        const uint32_t d = tmem_d + (uint32_t)((tap >> 1) * CH) + ((uint32_t)((tap & 1) * 16) << 16);
#pragma unroll
        for (int k = 0; k < 128 / 16; ++k)                     // 16 pixel rows = two 1024-byte row groups per step
          umma_bf16(d, a_d + (uint64_t)(k * 128), b_d + (uint64_t)(k * 128), IDESC, (tl | k) != 0);
        umma_commit(bar_empty + 8 * s);
        if (++tap == TAPS) { tap = 0; ++tl; }
      }
      umma_commit(bar_done);
    }
    __syncwarp();
  } else {
    // ============ producers: window gather -> blend -> swizzled col tile (as in the forward kernel) ============
    const int q = lane >> 3, l = lane & 7;
    const int wrow = warp >> 1, wcol = (warp & 1) * 8;
    const uint32_t offBase = sbase + Smem::OFFS_OFF + warp * NOB * OFF_WARP_BUF;
    const float* offF = reinterpret_cast<const float*>(smem + Smem::OFFS_OFF + warp * NOB * OFF_WARP_BUF);

    struct TileRef { const float* ob; const float* mb; uint32_t ok; };
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
      cp_col[k] = (i < NPLANES * PER_PLANE) ? (VEC_OFF ? e * 4 : e) : (1 << 28);
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
    TileRef pref = make_ref(0);
    int p_tl = 0, p_tap = 0, p_ring = 0;
    auto prefetch_next = [&]() {
      if (p_tl < my_tiles) {
        const uint32_t dst0 = offBase + p_ring * OFF_WARP_BUF;
#pragma unroll
        for (int k = 0; k < NCOPY; ++k) {
          if (pref.ok & (1u << k)) {
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
    // gout tile of a tile: row = pixel (the contraction index), 128 B = 64 cout, swizzled like the col tile
    auto fill_g = [&](int tl) {
      int n, ty0, tx0;
      tile_coords(tl, n, ty0, tx0);
      const __nv_bfloat16* gn = gout + (size_t)n * gs_n;
      const uint32_t gdst = sG + (tl & 1) * G_TILE;
      for (int i = tid; i < 128 * 8; i += CWARPS * 32) {
        const int r = i >> 3, c = i & 7;
        const int gy = ty0 + (r >> 4), gx = tx0 + (r & 15);
        const bool ok = gy < H && gx < W;
        const __nv_bfloat16* src = ok ? gn + ((size_t)gy * W + gx) * CH + c * 8 : gn;
        cp_async_16_zfill(gdst + sw128_offset(r, c * 16), src, ok);
      }
    };

    if (my_tiles > 0) {
      fill_g(0);
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_gfull);
    }
    prefetch_next();
    int it = 0, ring = 0, oring = 0;
    uint32_t ring_ph = 0;
    for (int tl = 0; tl < my_tiles; ++tl) {
      int n, ty0, tx0;
      tile_coords(tl, n, ty0, tx0);
      const __nv_bfloat16* xn = x + (size_t)n * xs_n;
      const int wy0 = ty0 - PAD, wx0 = tx0 - PAD;
      const uint32_t win = sWin + (tl & 1) * WIN_BYTES + l * 16;
      const int gy = ty0 + wrow;
      float pyb[2], pxb[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int gx = tx0 + wcol + j * 4 + q;
        pyb[j] = (gy < H && gx < W) ? (float)(gy - 1) : -100000.f;   // dead pixel: sample rejected -> zero row
        pxb[j] = (float)(gx - 1);
      }
      if (wy0 < 0 || wx0 < 0 || wy0 + WH > H || wx0 + WW > W) {
        uint8_t* wb = smem + Smem::WIN_OFF + (tl & 1) * WIN_BYTES;
        for (int i = tid; i < WH * WW * 8; i += CWARPS * 32) {
          const int cell = i >> 3;
          const int cy = wy0 + cell / WW, cx = wx0 + cell % WW;
          if ((unsigned)cy >= (unsigned)H || (unsigned)cx >= (unsigned)W)
            *reinterpret_cast<uint4*>(wb + i * 16) = make_uint4(0, 0, 0, 0);
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(CWARPS * 32) : "memory");
      }
      mbar_wait(bar_winf + 8 * (tl & 1), (tl >> 1) & 1);
      for (int tap = 0; tap < TAPS; ++tap, ++it) {
        // the gout tile of the next tile: issued once MMA(tap 0 of this tile) -- hence every MMA of tile
        // tl-1, the last reader of that buffer -- has retired (the empty wait of tap 3 below proves it)
        static_assert(NOB == 2, "one prefetch in flight: wait for everything");
        cp_async_wait<0>();                                    // offsets of `it` (and, at tap >= 5, the gout tile)
        if (tap == 6 && tl + 1 < my_tiles) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_gfull + 8 * ((tl + 1) & 1));
        }
        __syncwarp();
        prefetch_next();
        const float* so = offF + oring * (OFF_WARP_BUF / 4);
        if (++oring == NOB) oring = 0;
        const int ti = (tap * 11) >> 5, tj = tap - ti * 3;
        uint32_t res[2][4];
        constexpr int NI = 2 * NSUB, CW = 4 / NSUB;           // samples per thread and tap; words per corner
#pragma unroll
        for (int u = 0; u < NI; ++u) {
          const int j = u / NSUB, sb = u % NSUB;
          const int g = NSUB == 2 ? 2 * l + sb : (l * DG) / 8;
          const int col = (j * 4 + q) ^ (((g >> 2) & 1) << 2);
          // dead pixels (outside the image) have no staged offsets: whatever the buffer holds must not
          // reach the col tile, every row of which is summed into dW
          const bool live = pyb[j] > -50000.f;
          const float dy = live ? so[(0 * DG + g) * PLW + col] : 0.f;
          const float dx = live ? so[(1 * DG + g) * PLW + col] : 0.f;
          const float mk = live ? so[(2 * DG + g) * PLW + col] : 0.f;
          const float py = (pyb[j] + (float)ti) + dy;
          const float px = (pxb[j] + (float)tj) + dx;
          const int y0 = __float2int_rd(py), x0 = __float2int_rd(px);
          const float ly = py - (float)y0, lx = px - (float)x0;
          float wy0f = mk * (1.f - ly), wy1f = mk * ly, wx0f = 1.f - lx, wx1f = lx;
          const int ry = y0 - wy0, rx = x0 - wx0;
          uint32_t v[4][CW];
          if ((unsigned)ry < (unsigned)(WH - 1) && (unsigned)rx < (unsigned)(WW - 1)) {
            const uint32_t a00 = win + (uint32_t)(ry * WW + rx) * 128u + sb * 8u;
            const uint32_t ad[4] = {a00, a00 + 128u, a00 + WW * 128u, a00 + WW * 128u + 128u};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (CW == 4) {
                const uint4 t = lds128(ad[c]);
                v[c][0] = t.x; v[c][1] = t.y; v[c][CW - 2] = t.z; v[c][CW - 1] = t.w;
              } else {
                const uint2 t = lds64(ad[c]);
                v[c][0] = t.x; v[c][1] = t.y;
              }
            }
          } else {
            wy0f = ((unsigned)y0 < (unsigned)H) ? wy0f : 0.f;
            wy1f = ((unsigned)y0 + 1u < (unsigned)H) ? wy1f : 0.f;
            wx0f = ((unsigned)x0 < (unsigned)W) ? wx0f : 0.f;
            wx1f = ((unsigned)x0 + 1u < (unsigned)W) ? wx1f : 0.f;
            const int ys = min(max(y0, -1), H), xs = min(max(x0, -1), W);
            const int cy0 = min(max(ys, 0), H - 1), cy1 = min(max(ys + 1, 0), H - 1);
            const int cx0 = min(max(xs, 0), W - 1), cx1 = min(max(xs + 1, 0), W - 1);
            const uint32_t b00 = (uint32_t)(cy0 * W + cx0) * CH + l * 8 + sb * 4;
            const uint32_t sx = (uint32_t)(cx1 - cx0) * CH, sy = (uint32_t)((cy1 - cy0) * W) * CH;
            const uint32_t bo[4] = {b00, (uint32_t)(b00 + sx), (uint32_t)(b00 + sy), (uint32_t)(b00 + sy + sx)};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (CW == 4) {
                const uint4 t = __ldg(reinterpret_cast<const uint4*>(xn + bo[c]));
                v[c][0] = t.x; v[c][1] = t.y; v[c][CW - 2] = t.z; v[c][CW - 1] = t.w;
              } else {
                const uint2 t = __ldg(reinterpret_cast<const uint2*>(xn + bo[c]));
                v[c][0] = t.x; v[c][1] = t.y;
              }
            }
          }
          const float w00 = wy0f * wx0f, w01 = wy0f * wx1f, w10 = wy1f * wx0f, w11 = wy1f * wx1f;
#pragma unroll
          for (int e = 0; e < CW; ++e) {
            const float lo = w00 * bf16lo_to_f32(v[0][e]) + w01 * bf16lo_to_f32(v[1][e]) + w10 * bf16lo_to_f32(v[2][e]) +
                             w11 * bf16lo_to_f32(v[3][e]);
            const float hi = w00 * bf16hi_to_f32(v[0][e]) + w01 * bf16hi_to_f32(v[1][e]) + w10 * bf16hi_to_f32(v[2][e]) +
                             w11 * bf16hi_to_f32(v[3][e]);
            res[j][sb * CW + e] = pack_bf16x2(lo, hi);
          }
        }
        if (it >= NSA) mbar_wait(bar_empty + 8 * ring, ring_ph ^ 1);
        if (tap == 3 && tl + 1 < my_tiles) fill_g(tl + 1);     // joins the cp.async group committed by the next prefetch
        const uint32_t aStage = sA + ring * A_TILE;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int m = wrow * 16 + wcol + j * 4 + q;
          const uint32_t dst = aStage + sw128_offset(m, l * 16);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst), "r"(res[j][0]), "r"(res[j][1]),
                       "r"(res[j][2]), "r"(res[j][3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_full + 8 * ring);
          if (tap == TAPS - 1) mbar_arrive(bar_wine + 8 * (tl & 1));
        }
        if (++ring == NSA) { ring = 0; ring_ph ^= 1; }
      }
    }
    cp_async_wait<0>();

    // ---- reduce this CTA's nine 64x64 accumulators into gweight[cout][cin][tap] ----
    if (my_tiles > 0) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
      const int qd = warp & 3, cq = warp >> 2;
      const int cin = qd * 16 + (lane & 15), half = lane >> 4;
#pragma unroll 1
      for (int p = 0; p < (TAPS + 1) / 2; ++p) {
        uint32_t acc[16];
        tmem_ld_x16(tmem_d + ((uint32_t)(qd * 32) << 16) + p * CH + cq * 16, acc);
        tmem_ld_wait();
        const int tap = 2 * p + half;
        if (tap < TAPS) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int co = cq * 16 + e;
            atomicAdd(gweight + ((size_t)co * CH + cin) * TAPS + tap, __uint_as_float(acc[e]));
          }
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == CWARPS) tmem_dealloc<TMEM_COLS>(tmem_d);
}
}  // namespace wgt
}  // namespace bwd

// --------------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------------
size_t dcn_backward_tc_workspace() { return (size_t)bwd::TAPS * bwd::B_TILE; }

bool dcn_backward_tc_eligible(const int64_t* gs, const int64_t* xs, const int64_t* gxs, const DcnGeom& g, bool has_gx) {
  auto dense = [&](const int64_t* s) {
    return s[1] == 1 && s[3] == 64 && s[2] == (int64_t)g.W * 64 && s[0] >= (int64_t)g.H * g.W * 64 && s[0] % 8 == 0;
  };
  const bool cfg = g.Cin == 64 && g.Cout == 64 && g.KH == 3 && g.KW == 3 && g.SH == 1 && g.SW == 1 && g.PH == 1 &&
                   g.PW == 1 && g.DH == 1 && g.DW == 1 && g.G == 1 &&
                   (g.DG == 1 || g.DG == 2 || g.DG == 4 || g.DG == 8 || g.DG == 16);
  if (!cfg || (long long)g.H * g.W > (g.DG == 16 ? (1ll << 23) : (1ll << 24)) || (long long)g.N * g.DG * 18 * g.H * g.W >= (1ll << 40)) return false;
  if (!dense(gs) || !dense(xs)) return false;
  if (has_gx && !dense(gxs)) return false;
  return true;
}

template <int DG>
static int launch_bwd_tc(const void* gout, const int64_t* gs, const void* x, const int64_t* xs, const float* offset,
                         const float* mask, const void* weight, float* gx32, const int64_t* gxs, float* goffset,
                         float* gmask, float* gweight32, float* gbias32, const DcnGeom& g, void* workspace,
                         unsigned which, cudaStream_t st) {
  using namespace bwd;
  const int H = g.H, W = g.W;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const int tiles_per_img = tiles_x * tiles_y, total = tiles_per_img * g.N;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = total < sms ? total : sms;
  int rc = EAVSR_OK;
  if ((which & 1u) && (gx32 || goffset || gmask)) {
    dcn_pack_weight_t<<<(TAPS * CH * CH + 255) / 256, 256, 0, st>>>((const __nv_bfloat16*)weight, (uint8_t*)workspace);
    rc = check_launch("dcn_backward(pack)");
    if (rc) return rc;
    const size_t P = (size_t)H * W;
    if (gx32) cudaMemsetAsync(gx32, 0, (size_t)g.N * (size_t)gxs[0] * sizeof(float), st);
    if (DG < 8) {
      if (goffset) cudaMemsetAsync(goffset, 0, (size_t)g.N * DG * 18 * P * sizeof(float), st);
      if (gmask) cudaMemsetAsync(gmask, 0, (size_t)g.N * DG * 9 * P * sizeof(float), st);
    }
    auto k = data::dcn_bwd_data_kernel<DG>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, data::Smem::DYN);
    if (e != cudaSuccess) { set_error("dcn_backward(data): smem attr: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
    k<<<grid, THREADS, data::Smem::DYN, st>>>((const __nv_bfloat16*)gout, (const __nv_bfloat16*)x, offset, mask,
                                               (const uint8_t*)workspace, gx32, goffset, gmask, H, W, (long long)xs[0],
                                               (long long)gs[0], gx32 ? (long long)gxs[0] : 0ll, tiles_x,
                                               tiles_per_img, total);
    rc = check_launch("dcn_backward(data, tcgen05)");
    if (rc) return rc;
  }
  if ((which & 2u) && gweight32) {
    cudaMemsetAsync(gweight32, 0, (size_t)CH * CH * TAPS * sizeof(float), st);
    const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(offset) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(mask) & 15u) == 0);
    auto kv = wgt::dcn_bwd_weight_kernel<DG, true>;
    auto ks = wgt::dcn_bwd_weight_kernel<DG, false>;
    auto k = vec ? kv : ks;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, wgt::Cfg<DG>::DYN);
    if (e != cudaSuccess) { set_error("dcn_backward(weight): smem attr: %s", cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
    k<<<grid, THREADS, wgt::Cfg<DG>::DYN, st>>>((const __nv_bfloat16*)gout, (const __nv_bfloat16*)x, offset, mask,
                                              gweight32, H, W, (long long)xs[0], (long long)gs[0], tiles_x,
                                              tiles_per_img, total);
    rc = check_launch("dcn_backward(weight, tcgen05)");
    if (rc) return rc;
  }
  if (gbias32) {
    cudaMemsetAsync(gbias32, 0, CH * sizeof(float), st);
    const long long px = (long long)g.N * H * W;
    int blocks = (int)((px + 1023) / 1024);
    blocks = blocks < 1 ? 1 : (blocks > 2 * sms ? 2 * sms : blocks);
    dcn_bwd_bias_nhwc<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)gout, gbias32, g.N, (long long)H * W, (long long)gs[0]);
    rc = check_launch("dcn_backward(bias)");
  }
  return rc;
}

// which: bit 0 = data gradients, bit 1 = weight gradient on the tensor-core kernels; d(bias) always (NHWC kernel)
int dcn_backward_tc(const void* gout, const int64_t* gs, const void* x, const int64_t* xs, const float* offset,
                    const float* mask, const void* weight, float* gx32, const int64_t* gxs, float* goffset,
                    float* gmask, float* gweight32, float* gbias32, const DcnGeom& g, void* workspace, unsigned which,
                    cudaStream_t st) {
  switch (g.DG) {
    case 16: return launch_bwd_tc<16>(gout, gs, x, xs, offset, mask, weight, gx32, gxs, goffset, gmask, gweight32, gbias32, g, workspace, which, st);
    case 8: return launch_bwd_tc<8>(gout, gs, x, xs, offset, mask, weight, gx32, gxs, goffset, gmask, gweight32, gbias32, g, workspace, which, st);
    case 4: return launch_bwd_tc<4>(gout, gs, x, xs, offset, mask, weight, gx32, gxs, goffset, gmask, gweight32, gbias32, g, workspace, which, st);
    case 2: return launch_bwd_tc<2>(gout, gs, x, xs, offset, mask, weight, gx32, gxs, goffset, gmask, gweight32, gbias32, g, workspace, which, st);
    case 1: return launch_bwd_tc<1>(gout, gs, x, xs, offset, mask, weight, gx32, gxs, goffset, gmask, gweight32, gbias32, g, workspace, which, st);
  }
  set_error("dcn_backward: deform_groups %d not supported by the tensor-core path", g.DG);
  return EAVSR_ERR_INVALID;
}

}  // namespace eavsr
