// 3x3 convolution 64 -> 64 channels (NHWC bf16, stride 1, pad 1) as an implicit GEMM on
// tcgen05 / TMEM, with bias + (Leaky)ReLU and the channel sums of the channel-attention layer
// fused into the epilogue.  This is SURVEY.md section 8 row f3: the RCAB residual backbone
// (models/networks.py:449-482, models/eavsrp_model.py:366-400) is 65 % of EAVSR+'s FLOPs and,
// once the alignment path is native, 7 700 of these convolutions per 30-frame clip.
//
// D[128 positions x 64 cout] = sum_{tap} A_tap[128 x 64 cin] * W_tap[64 x 64]
//
// The trick that makes the A operands free: the CTA stages ONE halo tile of the input (6 rows x 32 columns of pixels)
// as a K-major, 128-byte-swizzled UMMA tile whose rows are the halo pixels (128 B = 64 channels per row, 16-byte
// chunk c of row p stored at chunk c ^ (p & 7)).  A shifted view of it -- the descriptor start address advanced by
// whole rows -- is again a valid operand (the 128B swizzle is a function of absolute address bits): no im2col, no
// re-staging, no second copy.  (The first version used the non-swizzled layout with 16-byte rows; every 8x16 B core
// matrix straddled two 128-byte lines and the MMA issue thread sat blocked on tcgen05.mma:
// profiles/r1_conv3x3_tc_ncu.txt.)  Round 1 issued one N = 64 MMA per tap and K step (36 per tile, views advanced by
// r*32 + s rows); since round 2 ONE N = 192 MMA covers the three taps of a kernel row (view advanced by r*32 rows,
// the three weight tiles of the row as one B operand) and the epilogue adds the three accumulator blocks one and
// two TMEM lanes apart (12 MMAs per tile).  One tile = 128 consecutive halo positions = a 4 x 32
// block of which 4 x 30 are real outputs (the 2 wrap-around columns per row are computed and dropped: 6 % waste).
//
// Warp roles (544 threads, 1 CTA / SM, persistent over tiles and over the layers of a chain):
//   warp  4    one lane issues the MMAs (B: 9 resident 128B-swizzled 64x64 weight tiles) and commits to mbarriers;
//              in tensor-load layers it also issues the halo tiles' 4-D tensor loads (6-stage ring: it re-requests a
//              stage when the MMAs that read it have completed)
//   warps 5-12 epilogue team 0, two groups of four: group g drains output channels 32g .. 32g+31:
//              tcgen05.ld x3 -> shifted sum -> +bias -> activation -> bf16 NHWC store (32-byte st.global.v8), channel
//              sums kept in registers across the CTA's tiles and flushed with one atomicAdd per channel
//   warps 0-3, 13-16
//              plain tensor-load layers: epilogue team 1 (the teams alternate tiles; a tile's epilogue is ~450
//              dependent instructions per warp and paced the pipeline, profiles/r2_conv3x3_tc_ncu.txt);
//              fused-input layers (channel attention y = res * scale + skip of the previous block folded in): the
//              eight warps transform the TMA-staged skip / res tiles in shared memory and write y out once;
//              without tensor maps (fallback): cp.async producers, one stage per warp, register-path transform.
// Two TMEM accumulators (2 x 192 columns) decouple the epilogues from the next tile's MMAs.
#include <cstdlib>

#include <cuda.h>

#include "common.cuh"

namespace eavsr {
namespace {

constexpr int CV_CH = 64;
constexpr int CV_TR = 4, CV_TC = 30;            // real outputs per tile
constexpr int CV_PW = 32;                       // halo pitch in pixels (= MMA rows per output row)
constexpr int CV_HROWS = (CV_TR + 2) * CV_PW;   // 192 staged halo positions
constexpr int CV_ROWS = 200;                    // + padding rows read by the wrap-around outputs
constexpr int CV_ASTAGE = CV_ROWS * 128;        // 200 rows of 64 bf16 (25 x 1024 B: stages stay atom-aligned)
constexpr int CV_NS = 6;                        // halo stages (plain layers: one TMA tensor load per tile and stage)
constexpr int CV_NSF = 4;                       // stages used when producer WARPS build the tile: one per warp
constexpr int CV_BTILE = CV_CH * CV_CH * 2;     // 8 KB per tap
constexpr int CV_PRODUCERS = 128;
// 4 producer + 1 MMA + 2 x 4 epilogue + 4 helper-producer warps.  17 warps put 5 on one SM sub-partition
// (16 K registers each), so the kernel is held to 96 registers per thread -- 116 would compile for a
// 64 K register file but fails to launch.
constexpr int CV_THREADS = 544;
constexpr int CV_ACC = 3 * CV_CH;              // accumulator columns per buffer: one 64-column block per tap column c
constexpr int CV_TMEM = 512;                    // two buffers of 192 (allocations are powers of two)
constexpr int CV_MAXN = 8;                      // images per call in the fused channel-attention mode

struct CvSmem {
  static constexpr int B_OFF = 0;                                  // 9 x 8 KB, 1024-aligned
  static constexpr int A_OFF = B_OFF + 9 * CV_BTILE;               // 3 stages
  static constexpr int BIAS_OFF = ((A_OFF + CV_NS * CV_ASTAGE + 15) / 16) * 16;   // 64 fp32
  static constexpr int RED_OFF = BIAS_OFF + CV_CH * 4;             // 64 fp32: per-CTA channel sums before the atomics
  static constexpr int SCALE_OFF = RED_OFF + 2 * CV_CH * 4;        // (one RED region per epilogue team)            // fused channel attention: CV_MAXN x 64 fp32 scales
  static constexpr int BAR_OFF = SCALE_OFF + CV_MAXN * CV_CH * 4;
  // two sets (layers alternate; the idle set is re-initialised off the critical path) of
  // full[NS], empty[NS], accf[2], acce[2], wbar; then the tmem slot
  // ... + lfull[NSF] (fused tensor-load layers: raw skip / res tiles landed) + rempty[2] (raw res buffer drained)
  static constexpr int NBARS = 2 * CV_NS + 5 + CV_NSF + 2;
  static constexpr int TOTAL = BAR_OFF + 2 * NBARS * 8 + 16;
  static constexpr int DYN = TOTAL + 1024;
};

__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}

__global__ void conv_pack_weight(const __nv_bfloat16* __restrict__ w, uint8_t* __restrict__ packed) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (tap, o, c)
  if (idx >= 9 * CV_CH * CV_CH) return;
  const int c = idx % CV_CH, o = (idx / CV_CH) % CV_CH, t = idx / (CV_CH * CV_CH);
  *reinterpret_cast<__nv_bfloat16*>(packed + (size_t)t * CV_BTILE + sw128_offset(o, c * 2)) =
      w[((size_t)o * CV_CH + c) * 9 + t];
}

// One convolution of a chain.  Fused input transform (the tail of the previous RCABlock, models/networks.py:
// 449-465): when `res` is set, the convolution's input is y = res * sigmoid(MLP(res_sums / HW)) + x, built by the
// producer warps while they stage the halo tile; the interior of every tile is also written to `y_out` (the next
// block's skip).
struct ConvLayerDev {
  const __nv_bfloat16* x;          // input (the skip tensor in the fused mode)
  const uint8_t* wpacked;
  const __nv_bfloat16* bias;
  __nv_bfloat16* out;
  float* chan_sums;                // (n, 64) or null
  const __nv_bfloat16* res;        // fused mode: residual branch of the previous block, or null
  const float* res_sums;           // (n, 64) channel sums of res
  const __nv_bfloat16 *w1, *b1, *w2, *b2;
  __nv_bfloat16* y_out;
  float slope;
  int tma;                         // halo tiles come by tensor loads: `tm` over x (and `tm2` over res in the fused mode)
  // (64 ch, W, H, N) bf16 view of x, box (64, 32, 6, 1), 128-byte swizzle: the box lands in a stage exactly as the
  // K-major SW128 operand layout the MMA descriptors describe, out-of-image pixels zero-filled (= the padding)
  alignas(64) CUtensorMap tm;
  alignas(64) CUtensorMap tm2;
};
// A whole residual group in ONE launch (RCAGroup = 61 chained 64->64 convolutions,
// models/networks.py:467-482): the persistent CTAs walk the layers, separated by a grid-wide barrier (every layer
// reads what all CTAs of the previous one wrote).  What a launch per convolution pays 61 times -- launch gap, TMEM
// allocation, a cold 72 KB weight load in front of the first MMA, pipeline drain -- is paid once or hidden: the next
// layer's weights stream in while the CTA waits at the barrier.  Kernel parameters are the layer table itself
// (<= 64 layers, 24 KB of the 32 KB parameter space with the tensor maps), so nothing has to be staged in device memory.
constexpr int CV_MAXL = 64;
struct ChainParams {
  int nlayers, H, W, tiles_x, tiles_per_img, total_tiles, nimg;
  float inv_hw;
  unsigned* sync;                  // grid-barrier counter, zero at launch (only read when nlayers > 1)
  ConvLayerDev L[CV_MAXL];
};

#ifdef EAVSR_CONV_TRACE   // development only: per-phase timestamps of CTA 0 (tools/abl_build.sh, tools/prof_chain.py)
__device__ unsigned long long g_conv_trace[CV_MAXL * 16];
#define CV_TRACE(li, slot) do { if (blockIdx.x == 0) g_conv_trace[(li) * 16 + (slot)] = clock64(); } while (0)
// per-tile stamps of CTA 0, layers 0..7, tiles 0..7: k = 0 loads landed (transform warp 0), 1 tile staged / transform
// done (MMA thread past `full`), 2 MMAs issued, 3 accumulator ready (epilogue warp 5 or 0), 4 accumulator drained, 5 tile stored
__device__ unsigned long long g_conv_trace2[8 * 8 * 8];
#define CV_TRACE2(li, tl, k) do { if (blockIdx.x == 0 && (li) < 8 && (tl) < 8) g_conv_trace2[((li) * 8 + (tl)) * 8 + (k)] = clock64(); } while (0)
#else
#define CV_TRACE(li, slot) do { } while (0)
#define CV_TRACE2(li, tl, k) do { } while (0)
#endif

__device__ __forceinline__ void mbar_inval(uint32_t bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(CV_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ ChainParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sB = base + CvSmem::B_OFF, sA = base + CvSmem::A_OFF, bars0 = base + CvSmem::BAR_OFF;
  const uint32_t tmem_slot_addr = bars0 + 2 * CvSmem::NBARS * 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + CvSmem::BAR_OFF + 2 * CvSmem::NBARS * 8);
  auto init_bar_set = [&](uint32_t b0, const ConvLayerDev& l) {
    // full: two warps of lanes (register path) or 8 transform warps (fused) / the TMA-issuing thread or one warp (plain)
    const uint32_t full_count = l.res ? (l.tma ? 8 : 64) : (l.tma ? 1 : 32);
    for (int s = 0; s < CV_NS; ++s) {
      mbar_init(b0 + 8 * s, full_count);
      mbar_init(b0 + (CV_NS + s) * 8, 1);                     // empty
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(b0 + (2 * CV_NS + b) * 8, 1);                 // accf
      mbar_init(b0 + (2 * CV_NS + 2 + b) * 8, 8);             // acce
    }
    mbar_init(b0 + (2 * CV_NS + 4) * 8, 1);                   // weights
    for (int s = 0; s < CV_NSF; ++s) mbar_init(b0 + (2 * CV_NS + 5 + s) * 8, 1);          // lfull
    for (int r = 0; r < 2; ++r) mbar_init(b0 + (2 * CV_NS + 5 + CV_NSF + r) * 8, 8);      // rempty
  };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = P.H, W = P.W, tiles_x = P.tiles_x, tiles_per_img = P.tiles_per_img;
  const int first = blockIdx.x;
  const int my_tiles = (P.total_tiles - first + (int)gridDim.x - 1) / (int)gridDim.x;

  // zero the padding rows of every stage once (they only feed dropped wrap-around outputs)
  for (int i = tid; i < CV_NS * (CV_ROWS - CV_HROWS) * 8; i += CV_THREADS) {
    const int s = i / ((CV_ROWS - CV_HROWS) * 8), r = i % ((CV_ROWS - CV_HROWS) * 8);
    *reinterpret_cast<uint4*>(smem + CvSmem::A_OFF + s * CV_ASTAGE + CV_HROWS * 128 + r * 16) = make_uint4(0, 0, 0, 0);
  }
  if (warp == 4) tmem_alloc<CV_TMEM>(tmem_slot_addr);
  uint32_t tmem_d = 0;

#pragma unroll 1
  for (int li = 0; li < P.nlayers; ++li) {
    const ConvLayerDev& Ly = P.L[li];
    const bool fused = Ly.res != nullptr;
    const bool tma = !fused && Ly.tma != 0;
    const bool ftma = fused && Ly.tma != 0;                     // fused input built from TMA-staged skip / res tiles
    const int NS = tma ? CV_NS : CV_NSF;                        // stages in use this layer
    const __nv_bfloat16* __restrict__ x = Ly.x;
    // ---------------- layer prologue ----------------
    // Layers alternate between two mbarrier sets: this layer's set was initialised while the previous layer ran.
    const uint32_t bars = bars0 + (li & 1) * CvSmem::NBARS * 8;
    const uint32_t bar_full = bars, bar_empty = bars + CV_NS * 8, bar_accf = bars + 2 * CV_NS * 8;
    const uint32_t bar_acce = bar_accf + 16, bar_w = bar_acce + 16;
    const uint32_t bar_lfull = bar_w + 8, bar_rempty = bar_lfull + CV_NSF * 8;
    if (li == 0) {
      if (tid == 0) {
        init_bar_set(bars, Ly);
        fence_mbar_init();
      }
      fence_proxy_async_smem();                                 // the zeroed padding rows, before any MMA reads them
    }
    // (A) every role has left layer li-1 (its global writes are issued, its MMAs have retired); layer 0: set-up
    __syncthreads();
    if (tid == 0) CV_TRACE(li, 0);
    if (li > 0 && tid == 0)                                     // grid barrier, arrive: this CTA's layer li-1 is out
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(P.sync) : "memory");
    // The 72 KB of packed weights start streaming before anything else (they do not depend on the other CTAs):
    // in a chain they land while this CTA waits at the grid barrier.
    if (warp == 4 && elect_one()) {
      mbar_arrive_expect_tx(bar_w, 9 * CV_BTILE);
      for (int t = 0; t < 9; ++t) bulk_g2s(sB + t * CV_BTILE, Ly.wpacked + (size_t)t * CV_BTILE, CV_BTILE, bar_w);
    }
    if (tid < CV_CH) reinterpret_cast<float*>(smem + CvSmem::BIAS_OFF)[tid] = Ly.bias ? __bfloat162float(Ly.bias[tid]) : 0.f;
    if (li > 0) {
      if (tid == 0) {                                           // grid barrier, wait: all CTAs finished layer li-1
        const unsigned want = (unsigned)li * gridDim.x;
        const long long t0 = clock64();
        while (ld_acquire_gpu(P.sync) < want) {
          if (clock64() - t0 > 4000000000ll) __trap();         // a lost CTA must fail, not hang the GPU
        }
      }
      __syncthreads();                                          // (C)
    }
    if (tid == 0) CV_TRACE(li, 1);
    if (tid == 32 && li + 1 < P.nlayers) {                      // the other set is quiescent: re-arm it for layer li+1
      const uint32_t nb = bars0 + ((li + 1) & 1) * CvSmem::NBARS * 8;
      if (li > 0)
        for (int b = 0; b < CvSmem::NBARS; ++b) mbar_inval(nb + 8 * b);
      init_bar_set(nb, P.L[li + 1]);
      fence_mbar_init();
    }

    // producer warp w streams the halo tile of tile tl (== w mod 4) into stage w
    auto issue_tile = [&](int tl) {
      const int tile = first + tl * (int)gridDim.x;
      const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
      const int y0 = (rem / tiles_x) * CV_TR - 1, x0 = (rem % tiles_x) * CV_TC - 1;
      const __nv_bfloat16* xn = x + (size_t)n * H * W * CV_CH;
      const uint32_t stage = sA + warp * CV_ASTAGE;
#pragma unroll 4
      for (int i = lane; i < CV_HROWS * 8; i += 32) {
        const int p = i >> 3, ch = i & 7;
        const int gy = y0 + (p >> 5), gx = x0 + (p & 31);
        const bool ok = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
        const __nv_bfloat16* src = ok ? xn + ((size_t)gy * W + gx) * CV_CH + ch * 8 : xn;
        cp_async_16_zfill(stage + p * 128 + ((ch ^ (p & 7)) << 4), src, ok);
      }
      cp_async_commit();
    };
    // First loads in flight before the set-up barrier -- tile 0 only: right after the grid barrier all 148 SMs
    // start at once, and four tiles per SM (14 MB) in one burst delay the first tile, the one the MMA warp is
    // waiting for, by ~1.5 us of L2 bandwidth (phase timestamps, tools/prof_chain.py).  Warps 1-3 start their
    // tiles when tile 0 has landed; those land while tile 0 is being multiplied.
    // Tensor-load layers: one thread requests the first NS tiles back to back (the TMA unit serves them in order, so
    // tile 0 still lands first) and keeps the ring full from the producer branch below.
    auto issue_tile_tma = [&](int tl) {
      const int tile = first + tl * (int)gridDim.x;
      const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
      const int y0 = (rem / tiles_x) * CV_TR - 1, x0 = (rem % tiles_x) * CV_TC - 1;
      const int s = tl % CV_NS;
      mbar_arrive_expect_tx(bar_full + 8 * s, CV_HROWS * 128);
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
              "r"(sA + s * CV_ASTAGE),
          "l"(reinterpret_cast<uint64_t>(&Ly.tm)), "r"(0), "r"(x0), "r"(y0), "r"(n), "r"(bar_full + 8 * s)
          : "memory");
    };
    // Fused tensor-load layers: the skip tile lands in A stage tl % 4 (operand layout), the residual tile in one of
    // two raw buffers (A stages 4 and 5, same swizzle, so y = res * scale + skip is an address-wise update).
    auto issue_tile_fused = [&](int tl) {
      const int tile = first + tl * (int)gridDim.x;
      const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
      const int y0 = (rem / tiles_x) * CV_TR - 1, x0 = (rem % tiles_x) * CV_TC - 1;
      const int s = tl % CV_NSF, r = tl & 1;
      const uint32_t bar = bar_lfull + 8 * s;
      mbar_arrive_expect_tx(bar, 2 * CV_HROWS * 128);
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
              "r"(sA + s * CV_ASTAGE),
          "l"(reinterpret_cast<uint64_t>(&Ly.tm)), "r"(0), "r"(x0), "r"(y0), "r"(n), "r"(bar)
          : "memory");
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
              "r"(sA + (CV_NSF + r) * CV_ASTAGE),
          "l"(reinterpret_cast<uint64_t>(&Ly.tm2)), "r"(0), "r"(x0), "r"(y0), "r"(n), "r"(bar)
          : "memory");
    };
    if (ftma) {
      if (warp == 4 && elect_one()) {
        if (li > 0) asm volatile("fence.proxy.async.global;\n" ::: "memory");
        for (int tl = 0; tl < my_tiles && tl < 2; ++tl) issue_tile_fused(tl);
      }
    } else if (tma) {
      if (warp == 4 && elect_one()) {
        // (reader side of the cross-proxy hand-over: the acquire of the grid barrier was a generic-proxy operation)
        if (li > 0) asm volatile("fence.proxy.async.global;\n" ::: "memory");
        for (int tl = 0; tl < my_tiles && tl < CV_NS; ++tl) issue_tile_tma(tl);
      }
    } else if (!fused && warp == 0 && my_tiles > 0) issue_tile(0);
    if (fused && warp >= 5 && warp < 13) {
      // squeeze-excite MLP of the previous block, once per CTA: 256 threads = 4 hidden units x 64 channels
      float* scale = reinterpret_cast<float*>(smem + CvSmem::SCALE_OFF);
      float* part = reinterpret_cast<float*>(smem + CvSmem::RED_OFF);      // (free until the first flush)
      const int e = tid - 5 * 32, r = e >> 6, i = e & 63;
      for (int img = 0; img < P.nimg; ++img) {
        float pr = __bfloat162float(Ly.w1[r * CV_CH + i]) * (__ldcg(Ly.res_sums + img * CV_CH + i) * P.inv_hw);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) pr += __shfl_xor_sync(0xffffffffu, pr, off);
        if ((e & 31) == 0) part[e >> 5] = pr;
        asm volatile("bar.sync 2, 256;\n" ::: "memory");
        if (e < CV_CH) {
          float a = __bfloat162float(Ly.b2[e]);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            a += __bfloat162float(Ly.w2[e * 4 + j]) * fmaxf(__bfloat162float(Ly.b1[j]) + part[2 * j] + part[2 * j + 1], 0.f);
          scale[img * CV_CH + e] = 1.f / (1.f + __expf(-a));
        }
        asm volatile("bar.sync 2, 256;\n" ::: "memory");
      }
    }
    tc_fence_before();
    __syncthreads();                                            // (D) bias / scales / re-armed barriers are visible
    tc_fence_after();
    tmem_d = *tmem_slot;
    if (tid == 0) CV_TRACE(li, 2);

    if (ftma && (warp < 4 || warp >= 13)) {
      // ===================== fused input, tensor-load form: y = res * scale + skip, shared memory -> shared memory =====================
      // The register path below is bound by the round trips of its global loads (first tile 5 us after the set-up,
      // 1.56 us per tile: profiles/r2_conv_chain_trace.txt).  Here the TMA unit brings both tiles, two tiles ahead,
      // and the eight warps only transform: lane l owns channel chunk l & 7 (its 8 scales stay in registers) of rows
      // 24 w + (l >> 3) + 4 k; a quarter warp covers one 128-byte row, so every access is conflict-free.
      const int widx = warp >= 13 ? warp - 9 : warp;            // 0..7
      const int ch = lane & 7;
      const float* scale = reinterpret_cast<const float*>(smem + CvSmem::SCALE_OFF);
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int s = tl % CV_NSF, r = tl & 1;
        const int tile = first + tl * (int)gridDim.x;
        const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
        const int y0 = (rem / tiles_x) * CV_TR - 1, x0 = (rem % tiles_x) * CV_TC - 1;
        const size_t img = (size_t)n * H * W * CV_CH;
        float s8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s8[e] = scale[n * CV_CH + ch * 8 + e];
        mbar_wait(bar_lfull + 8 * s, (tl / CV_NSF) & 1);
        if (warp == 0 && lane == 0) CV_TRACE2(li, tl, 0);
        const uint32_t aS = sA + s * CV_ASTAGE, rS = sA + (CV_NSF + r) * CV_ASTAGE;
        constexpr int NB = CV_HROWS / 8 / 4;                    // 6 rows per lane
        uint4 rv[NB], sv[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int p = widx * (CV_HROWS / 8) + (lane >> 3) + 4 * b;
          const uint32_t o = (uint32_t)p * 128u + (uint32_t)((ch ^ (p & 7)) << 4);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n"
                       : "=r"(sv[b].x), "=r"(sv[b].y), "=r"(sv[b].z), "=r"(sv[b].w) : "r"(aS + o));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n"
                       : "=r"(rv[b].x), "=r"(rv[b].y), "=r"(rv[b].z), "=r"(rv[b].w) : "r"(rS + o));
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int p = widx * (CV_HROWS / 8) + (lane >> 3) + 4 * b;
          const uint32_t o = (uint32_t)p * 128u + (uint32_t)((ch ^ (p & 7)) << 4);
          const int hr = p >> 5, hc = p & 31;
          const int gy = y0 + hr, gx = x0 + hc;
          const uint32_t rw[4] = {rv[b].x, rv[b].y, rv[b].z, rv[b].w}, sw[4] = {sv[b].x, sv[b].y, sv[b].z, sv[b].w};
          uint32_t ow[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            ow[e] = pack_bf16x2(bf16lo_to_f32(rw[e]) * s8[2 * e] + bf16lo_to_f32(sw[e]),
                                bf16hi_to_f32(rw[e]) * s8[2 * e + 1] + bf16hi_to_f32(sw[e]));
          // (outside the image both tiles were zero-filled: y = 0 = the convolution's padding)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(aS + o), "r"(ow[0]), "r"(ow[1]), "r"(ow[2]), "r"(ow[3])
                       : "memory");
          if (hr >= 1 && hr <= CV_TR && hc >= 1 && hc <= CV_TC && gy < H && gx < W)      // this tile owns the pixel
            *reinterpret_cast<uint4*>(Ly.y_out + img + ((size_t)gy * W + gx) * CV_CH + ch * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_full + 8 * s);
          mbar_arrive(bar_rempty + 8 * r);
        }
      }
    } else if (!tma && (warp < 4 || warp >= 13)) {
      // ===================== producers =====================
      // (tensor-load layers have no producer warps: the MMA thread re-issues a stage's load when the MMAs that read
      // it have completed, and warps 0-3 / 13-16 form a second epilogue team)
      // (warps 13-16 only work in the fused-input mode, where they build the second half of the halo tile of
      // "their" stage: the transform is load-latency bound, two warps per stage keep enough loads in flight)
      const int phalf = warp >= 13 ? 1 : 0;
      const int warp_s = warp >= 13 ? warp - 13 : warp;         // stage / tile phase served by this warp
      // Warp w owns stage w and the tiles tl == w (mod 4): it waits for its stage to drain, streams the
      // halo tile in, waits for ITS copies only and publishes.  The four warps are independent, so up
      // to four tiles are in flight and a late MMA never delays the publication of a landed tile.
      const float* scale = reinterpret_cast<const float*>(smem + CvSmem::SCALE_OFF);
      for (int tl = (phalf && !fused) ? my_tiles : warp_s; tl < my_tiles; tl += CV_NSF) {
        const int u = tl / CV_NSF;
        if (u >= 1) mbar_wait(bar_empty + 8 * warp_s, (u - 1) & 1);
        if (!fused) {
          if (u == 0 && warp_s > 0) mbar_wait(bar_full, 0);      // tile 0 first (see the prologue)
          if (u >= 1 || warp_s > 0) issue_tile(tl);            // (tile 0 was issued in the prologue)
          cp_async_wait<0>();
        } else {
          const int tile = first + tl * (int)gridDim.x;
          const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
          const int y0 = (rem / tiles_x) * CV_TR - 1, x0 = (rem % tiles_x) * CV_TC - 1;
          const size_t img = (size_t)n * H * W * CV_CH;
          const float* sc = scale + n * CV_CH + (lane & 7) * 8;      // this lane always handles chunk lane & 7
          float s8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) s8[e] = sc[e];
          // batches of 8 chunks per lane: 16 independent 16-byte loads in flight before the first use
          constexpr int NB = 8;
          static_assert((CV_HROWS * 4) % (32 * NB) == 0, "whole batches per half tile");
#pragma unroll 1
          for (int i0 = lane + phalf * (CV_HROWS * 4); i0 < (phalf + 1) * (CV_HROWS * 4); i0 += 32 * NB) {
            uint4 rv[NB], sv[NB];
            uint32_t off[NB];                                   // element offsets (n <= 8 images of < 2^24 pixels)
            bool okv[NB];
            // volatile asm: ptxas otherwise sinks every load pair down to its use (one round trip per chunk).
            // .cg (L2 only): in a chain these buffers were written by other SMs earlier in this same launch and
            // are recycled from layer to layer, so a line this SM cached two layers ago must not be hit.
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const int p = (i0 + 32 * b) >> 3;
              const int gy = y0 + (p >> 5), gx = x0 + (p & 31);
              okv[b] = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
              off[b] = okv[b] ? (uint32_t)(img + ((size_t)gy * W + gx) * CV_CH) + (lane & 7) * 8 : 0u;
              asm volatile("ld.global.cg.v4.b32 {%0, %1, %2, %3}, [%4];\n"
                           : "=r"(rv[b].x), "=r"(rv[b].y), "=r"(rv[b].z), "=r"(rv[b].w) : "l"(Ly.res + off[b]));
              asm volatile("ld.global.cg.v4.b32 {%0, %1, %2, %3}, [%4];\n"
                           : "=r"(sv[b].x), "=r"(sv[b].y), "=r"(sv[b].z), "=r"(sv[b].w) : "l"(x + off[b]));
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const int p = (i0 + 32 * b) >> 3, ch = lane & 7;
              const int hr = p >> 5, hc = p & 31;
              const uint32_t rw[4] = {rv[b].x, rv[b].y, rv[b].z, rv[b].w}, sw[4] = {sv[b].x, sv[b].y, sv[b].z, sv[b].w};
              uint32_t ow[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                ow[e] = pack_bf16x2(bf16lo_to_f32(rw[e]) * s8[2 * e] + bf16lo_to_f32(sw[e]),
                                    bf16hi_to_f32(rw[e]) * s8[2 * e + 1] + bf16hi_to_f32(sw[e]));
              const uint4 yv = okv[b] ? make_uint4(ow[0], ow[1], ow[2], ow[3]) : make_uint4(0, 0, 0, 0);
              if (okv[b] && hr >= 1 && hr <= CV_TR && hc >= 1 && hc <= CV_TC)      // this tile owns the pixel
                *reinterpret_cast<uint4*>(Ly.y_out + off[b]) = yv;
              *reinterpret_cast<uint4*>(smem + CvSmem::A_OFF + warp_s * CV_ASTAGE + p * 128 + ((ch ^ (p & 7)) << 4)) = yv;
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        mbar_arrive(bar_full + 8 * warp_s);
      }
    } else if (warp == 4) {
      // ===================== MMA issuer =====================
      if (elect_one()) {
        mbar_wait(bar_w, 0);
        CV_TRACE(li, 3);
        // One MMA covers the three taps of a kernel ROW: the A operand is the halo tile advanced by r whole halo
        // rows (32 pixel rows = 4 swizzle atoms, so every view is atom-aligned) and the B operand is the three
        // 64 x 64 weight tiles of taps (r, 0..2) read as one N = 192 tile (they are contiguous in shared memory).
        // Accumulator block c then holds D_c[m] = sum_r halo(32 r + m) . W(r, c), and the output at halo position p
        // is D_0[p] + D_1[p + 1] + D_2[p + 2] -- the epilogue's job (a tile row is one 32-lane TMEM quadrant and only
        // its first 30 positions are outputs, so p + 2 never leaves the warp).  Per K = 16 step an SS-mode MMA now
        // fetches A 4 KB + B 6 KB for 96 cycles of tensor work instead of 3 x (4 + 2) KB for 3 x 32: the operand
        // fetch that bounded the N = 64 form (67 cycles per MMA, 36 per tile) is off the critical path.
        constexpr uint32_t IDESC = umma_idesc_bf16(128, CV_ACC);
        const uint64_t b_base = umma_desc_sw128_kmajor(sB);
        for (int tl = 0; tl < my_tiles; ++tl) {
          const int s = tl % NS, buf = tl & 1;
          if (tl >= 2) mbar_wait(bar_acce + 8 * buf, ((tl >> 1) - 1) & 1);
          mbar_wait(bar_full + 8 * s, (tl / NS) & 1);
          if (tl == 0) CV_TRACE(li, 4);
          if (tl == 1) CV_TRACE(li, 5);
          if (tl == my_tiles - 1) CV_TRACE(li, 6);
          tc_fence_after();
          CV_TRACE2(li, tl, 1);
          // fused tensor-load layers: request tile tl+2 now, BEFORE this tile's MMAs (their issue blocks while the
          // tensor pipe still works on tile tl-1, and the raw ring is only two deep).  Its raw buffer was drained by
          // the transform of tile tl (just waited for), its A stage by the MMAs of tile tl-2.
          if (ftma && tl + 2 < my_tiles) {
            const int nt = tl + 2;
            if (nt >= CV_NSF) mbar_wait(bar_empty + 8 * (nt % CV_NSF), ((nt / CV_NSF) - 1) & 1);
            mbar_wait(bar_rempty + 8 * (nt & 1), ((nt >> 1) - 1) & 1);
            issue_tile_fused(nt);
          }
          const uint64_t a_base = umma_desc_sw128_kmajor(sA + s * CV_ASTAGE);
          const uint32_t d = tmem_d + buf * CV_ACC;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // the swizzle XOR is taken from the absolute shared-memory address bits, so a shifted view needs no
              // base_offset (measured in round 1 on the per-tap views)
              const uint64_t adesc = a_base + (uint64_t)((r * CV_PW * 128) >> 4) + 2 * k;
              const uint64_t bdesc = b_base + (uint64_t)((r * 3 * CV_BTILE) >> 4) + 2 * k;
              umma_bf16(d, adesc, bdesc, IDESC, (r | k) != 0);
            }
          }
          umma_commit(bar_empty + 8 * s);
          umma_commit(bar_accf + 8 * buf);
          CV_TRACE2(li, tl, 2);
          if (tl == my_tiles - 1) CV_TRACE(li, 7);
          // tensor-load layers: refill the stage tile tl-1 used.  Its MMAs complete while tile tl's execute, so this
          // wait returns long before the tensor pipe runs dry and the ring stays CV_NS - 1 tiles ahead.
          if (tma && tl >= 1 && tl - 1 + CV_NS < my_tiles) {
            mbar_wait(bar_empty + 8 * ((tl - 1) % CV_NS), ((tl - 1) / CV_NS) & 1);
            issue_tile_tma(tl - 1 + CV_NS);
          }

        }
      }
      __syncwarp();
    } else {
      // ===================== epilogue =====================
      // Team 0 = warps 5..12 (TMEM lane quadrants 1,2,3,0, 1,2,3,0).  A tile's epilogue is ~450 dependent instructions
      // of ONE warp per (quadrant, channel half) -- 1.5 us, the step that paced the whole pipeline (the MMA thread sat
      // waiting for a free accumulator; ncu: profiles/r2_conv3x3_tc_ncu.txt) -- so in tensor-load layers the idle
      // producer warps 0..3 and 13..16 (quadrants 0,1,2,3, 1,2,3,0) form team 1 and the teams alternate tiles: team t
      // always drains accumulator buffer t.
      const int team = (warp >= 5 && warp < 13) ? 0 : 1;
      const int nteams = tma ? 2 : 1;
      const int q = warp & 3;                       // output row of the tile handled by this warp
      const int chalf = team == 0 ? (warp - 5) >> 2 : (warp >= 13 ? 1 : 0);   // this warp's 32 output channels
      constexpr int HC = CV_CH / 2;
      float* const chan_sums = Ly.chan_sums;
      __nv_bfloat16* const out = Ly.out;
      const float slope = Ly.slope;
      float csum[HC];                               // this thread's running channel sums (its pixels)
#pragma unroll
      for (int c = 0; c < HC; ++c) csum[c] = 0.f;
      const float* bsm = reinterpret_cast<const float*>(smem + CvSmem::BIAS_OFF) + chalf * HC;
      int cur_n = -1;
      auto flush = [&]() {
        if (chan_sums && cur_n >= 0) {
          // transpose-reduce over the 32 lanes: 31 shuffles leave lane l with the sum of channel l of this half
          // (step k adds (16 >> k) to the channel base when lane bit (16 >> k) is set: base = lane)
#pragma unroll
          for (int step = 0; step < 5; ++step) {
            const int m = 16 >> step, len = 16 >> step;
            const bool up = (lane & m) != 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (i < len) {
                const float send = up ? csum[i] : csum[i + len];
                const float keep = up ? csum[i + len] : csum[i];
                csum[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
              }
            }
          }
          // four warps hold the same 32 channels: combine them in shared memory, one global atomic per
          // channel and CTA (148 instead of 592 reductions on each of the 64 addresses)
          float* red = reinterpret_cast<float*>(smem + CvSmem::RED_OFF) + team * CV_CH;
          const int bar_id = 2 + team;                         // the team's own named barrier
          asm volatile("bar.sync %0, 256;\n" ::"r"(bar_id) : "memory");       // previous flush fully drained
          if (q == 0) red[chalf * HC + lane] = csum[0];
          asm volatile("bar.sync %0, 256;\n" ::"r"(bar_id) : "memory");
          if (q != 0) atomicAdd(red + chalf * HC + lane, csum[0]);
          asm volatile("bar.sync %0, 256;\n" ::"r"(bar_id) : "memory");
          if (q == 0) atomicAdd(chan_sums + cur_n * CV_CH + chalf * HC + lane, red[chalf * HC + lane]);
        }
#pragma unroll
        for (int c = 0; c < HC; ++c) csum[c] = 0.f;
      };
      const float rcp_tpi = 1.f / (float)tiles_per_img, rcp_tx = 1.f / (float)tiles_x;
      auto divmod = [](int a, int d, float rcp, int& qo, int& r) {   // exact for a < 2^22: float estimate + one correction
        qo = __float2int_rz((float)a * rcp);
        r = a - qo * d;
        if (r < 0) { r += d; --qo; }
        if (r >= d) { r -= d; ++qo; }
      };
      for (int tl = team; tl < my_tiles; tl += nteams) {
        const int buf = tl & 1;
        const int tile = first + tl * (int)gridDim.x;
        int n, rem, ty, tx;
        divmod(tile, tiles_per_img, rcp_tpi, n, rem);
        divmod(rem, tiles_x, rcp_tx, ty, tx);
        if (n != cur_n) { flush(); cur_n = n; }
        const int oy = ty * CV_TR + q, ox = tx * CV_TC + lane;
        const bool valid = lane < CV_TC && oy < H && ox < W;
        mbar_wait(bar_accf + 8 * buf, (tl >> 1) & 1);
        if ((warp == 5 || warp == 0) && lane == 0) { if (tl == 0) CV_TRACE(li, 8); if (tl == my_tiles - 1) CV_TRACE(li, 9); CV_TRACE2(li, tl, 3); }
        tc_fence_after();
        // out(p) = D_0(p) + D_1(p + 1) + D_2(p + 2): three accumulator blocks, the second and third read one and two
        // TMEM lanes (= lanes of this warp) further.  16 channels (one 32-byte sector of the NHWC row) at a time.
        const uint32_t tbase = tmem_d + ((uint32_t)(q * 32) << 16) + buf * CV_ACC + chalf * HC;
        __nv_bfloat16* op = out + ((size_t)n * H * W + (size_t)oy * W + ox) * CV_CH + chalf * HC;
#pragma unroll
        for (int j = 0; j < HC; j += 16) {
          uint32_t a0[16], a1[16], a2[16];
          tmem_ld_32x16(tbase + j, a0);
          tmem_ld_32x16(tbase + CV_CH + j, a1);
          tmem_ld_32x16(tbase + 2 * CV_CH + j, a2);
          tmem_ld_wait();
          if (j + 16 == HC) {                                  // accumulator drained: the MMAs of tile tl + 2 may start
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acce + 8 * buf);
            if ((warp == 5 || warp == 0) && lane == 0) CV_TRACE2(li, tl, 4);
          }
          float f[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float t = __uint_as_float(a0[c]) + __shfl_down_sync(0xffffffffu, __uint_as_float(a1[c]), 1) +
                            __shfl_down_sync(0xffffffffu, __uint_as_float(a2[c]), 2) + bsm[j + c];
            f[c] = valid ? (t > 0.f ? t : t * slope) : 0.f;
          }
          if (valid) {                                         // one full 32-byte sector per store
            uint32_t u[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) u[e] = pack_bf16x2(f[2 * e], f[2 * e + 1]);
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(op + j), "r"(u[0]), "r"(u[1]),
                         "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
                         : "memory");
          }
          if (chan_sums) {
#pragma unroll
            for (int c = 0; c < 16; ++c) csum[j + c] += f[c];
          }
        }
        if ((warp == 5 || warp == 0) && lane == 0) CV_TRACE2(li, tl, 5);
      }
      flush();
      if (warp == 5 && lane == 0) CV_TRACE(li, 10);
    }
    // What this thread stored (outputs, y_out) may be read by another CTA's TENSOR loads in the next layer: order the
    // generic-proxy writes before the async proxy, ahead of the barrier (A) / the grid-barrier arrive that publish them.
    if (li + 1 < P.nlayers) asm volatile("fence.proxy.async.global;\n" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<CV_TMEM>(tmem_d);
}

}  // namespace
}  // namespace eavsr

using namespace eavsr;

extern "C" size_t eavsr_conv3x3_packed_weight_bytes(void) { return (size_t)9 * CV_BTILE; }

extern "C" int eavsr_conv3x3_pack_weight(const void* weight, void* packed, int cin, int cout, int dtype,
                                         void* stream) {
  EAVSR_REQUIRE(weight && packed, "conv3x3_pack_weight: null pointer");
  if (cin != CV_CH || cout != CV_CH || dtype != EAVSR_BF16) {
    set_error("conv3x3: only 64->64 bf16 is implemented (got %d->%d, dtype %d)", cin, cout, dtype);
    return EAVSR_ERR_UNSUPPORTED;
  }
  conv_pack_weight<<<(9 * CV_CH * CV_CH + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)weight,
                                                                                   (uint8_t*)packed);
  return check_launch("conv3x3_pack_weight");
}

namespace eavsr {
namespace {
// Common launch path: one layer = an ordinary launch, several = a cooperative launch (all CTAs must be resident
// for the grid barrier between layers).
int conv3x3_launch_chain(ChainParams& P, int n, int h, int w, int dtype, cudaStream_t st, const char* who) {
  EAVSR_REQUIRE(n > 0 && h > 0 && w > 0, "%s: empty tensor", who);
  EAVSR_REQUIRE(P.nlayers >= 1 && P.nlayers <= CV_MAXL, "%s: 1..%d layers per launch (got %d)", who, CV_MAXL, P.nlayers);
  if (dtype != EAVSR_BF16) {
    set_error("conv3x3: only 64->64 bf16 is implemented (dtype %d)", dtype);
    return EAVSR_ERR_UNSUPPORTED;
  }
  bool any_fused = false;
  for (int i = 0; i < P.nlayers; ++i) {
    const ConvLayerDev& L = P.L[i];
    EAVSR_REQUIRE(L.x && L.wpacked && L.out, "%s: null pointer (layer %d)", who, i);
    EAVSR_REQUIRE(((reinterpret_cast<uintptr_t>(L.x) | reinterpret_cast<uintptr_t>(L.out) |
                    reinterpret_cast<uintptr_t>(L.wpacked)) & 15u) == 0,
                  "%s: x / out / packed weights must be 16-byte aligned dense NHWC (layer %d)", who, i);
    if (L.res) {
      any_fused = true;
      EAVSR_REQUIRE(L.res_sums && L.w1 && L.b1 && L.w2 && L.b2 && L.y_out, "%s: null pointer (fused layer %d)", who, i);
      EAVSR_REQUIRE(((reinterpret_cast<uintptr_t>(L.res) | reinterpret_cast<uintptr_t>(L.y_out)) & 15u) == 0,
                    "%s: res / y_out must be 16-byte aligned dense NHWC (layer %d)", who, i);
    }
  }
  if (any_fused && n > CV_MAXN) {
    set_error("%s: at most %d images per call with a fused channel-attention input (got %d)", who, CV_MAXN, n);
    return EAVSR_ERR_UNSUPPORTED;
  }
  const int tiles_x = ceil_div(w, CV_TC), tiles_y = ceil_div(h, CV_TR);
  const int tiles_per_img = tiles_x * tiles_y;
  const long long total = (long long)tiles_per_img * n;
  EAVSR_REQUIRE(total < (1ll << 22), "%s: too many tiles", who);   // (float-reciprocal tile coordinates are exact below 2^22)
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(total < sms ? total : sms);
  P.H = h; P.W = w; P.tiles_x = tiles_x; P.tiles_per_img = tiles_per_img; P.total_tiles = (int)total; P.nimg = n;
  P.inv_hw = 1.f / ((float)h * (float)w);
  for (int i = 0; i < P.nlayers; ++i) {
    ConvLayerDev& L = P.L[i];
    L.tma = 0;
    const unsigned long long dims[4] = {CV_CH, (unsigned long long)w, (unsigned long long)h, (unsigned long long)n};
    const unsigned long long str[3] = {CV_CH * 2ull, (unsigned long long)w * CV_CH * 2, (unsigned long long)h * w * CV_CH * 2};
    const unsigned box[4] = {CV_CH, CV_PW, CV_TR + 2, 1};
    // development switches (A/B timing): EAVSR_CONV_NO_TMA=1 keeps every layer on the cp.async / register paths,
    // EAVSR_CONV_NO_FTMA=1 only the fused-input layers
    static const bool no_tma = getenv("EAVSR_CONV_NO_TMA") != nullptr, no_ftma = getenv("EAVSR_CONV_NO_FTMA") != nullptr;
    if (no_tma || (L.res && no_ftma)) continue;
    if (encode_tensor_map(&L.tm, EAVSR_BF16, 4, L.x, dims, str, box, 3 /* CU_TENSOR_MAP_SWIZZLE_128B */) &&
        (!L.res || encode_tensor_map(&L.tm2, EAVSR_BF16, 4, L.res, dims, str, box, 3)))
      L.tma = 1;
  }
  cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CvSmem::DYN);
  if (e != cudaSuccess) { set_error("%s: smem attr: %s", who, cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
  if (P.nlayers == 1) {
    P.sync = nullptr;
    conv3x3_tc_kernel<<<grid, CV_THREADS, CvSmem::DYN, st>>>(P);
    return check_launch(who);
  }
  EAVSR_REQUIRE(P.sync && (reinterpret_cast<uintptr_t>(P.sync) & 3u) == 0, "%s: a chain needs a 4-byte sync workspace", who);
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop) { set_error("%s: device has no cooperative launch", who); return EAVSR_ERR_UNSUPPORTED; }
  e = cudaMemsetAsync(P.sync, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) { set_error("%s: memset: %s", who, cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
  void* args[] = {&P};
  e = cudaLaunchCooperativeKernel((const void*)conv3x3_tc_kernel, dim3(grid), dim3(CV_THREADS), args, CvSmem::DYN, st);
  if (e != cudaSuccess) { cudaGetLastError(); set_error("%s: cooperative launch: %s", who, cudaGetErrorString(e)); return EAVSR_ERR_CUDA; }
  return check_launch(who);
}
}  // namespace
}  // namespace eavsr

extern "C" int eavsr_conv3x3_forward(const void* x, const void* packed_weight, const void* bias, void* out,
                                     float* channel_sums, int n, int cin, int cout, int h, int w,
                                     float negative_slope, int dtype, unsigned flags, void* stream) {
  if (cin != CV_CH || cout != CV_CH) {
    set_error("conv3x3: only 64->64 bf16 is implemented (got %d->%d, dtype %d)", cin, cout, dtype);
    return EAVSR_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (channel_sums && !(flags & EAVSR_CONV_SUMS_PREZEROED) && n > 0) {
    cudaError_t em = cudaMemsetAsync(channel_sums, 0, (size_t)n * CV_CH * sizeof(float), st);
    if (em != cudaSuccess) { set_error("conv3x3_forward: memset: %s", cudaGetErrorString(em)); return EAVSR_ERR_CUDA; }
  }
  static thread_local ChainParams P;
  P.nlayers = 1;
  P.L[0] = ConvLayerDev{(const __nv_bfloat16*)x, (const uint8_t*)packed_weight, (const __nv_bfloat16*)bias,
                        (__nv_bfloat16*)out, channel_sums, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                        negative_slope, 0};
  return conv3x3_launch_chain(P, n, h, w, dtype, st, "conv3x3_forward");
}

extern "C" int eavsr_conv3x3_ca_forward(const void* skip, const void* res, const float* res_sums, const void* w1,
                                        const void* b1, const void* w2, const void* b2, void* y_out,
                                        const void* packed_weight, const void* bias, void* out, float* channel_sums,
                                        int n, int h, int w, float negative_slope, int dtype, unsigned flags,
                                        void* stream) {
  EAVSR_REQUIRE(skip && res && res_sums && w1 && b1 && w2 && b2 && y_out, "conv3x3_ca_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (channel_sums && !(flags & EAVSR_CONV_SUMS_PREZEROED) && n > 0) {
    cudaError_t em = cudaMemsetAsync(channel_sums, 0, (size_t)n * CV_CH * sizeof(float), st);
    if (em != cudaSuccess) { set_error("conv3x3_ca_forward: memset: %s", cudaGetErrorString(em)); return EAVSR_ERR_CUDA; }
  }
  static thread_local ChainParams P;
  P.nlayers = 1;
  P.L[0] = ConvLayerDev{(const __nv_bfloat16*)skip, (const uint8_t*)packed_weight, (const __nv_bfloat16*)bias,
                        (__nv_bfloat16*)out, channel_sums, (const __nv_bfloat16*)res, res_sums,
                        (const __nv_bfloat16*)w1, (const __nv_bfloat16*)b1, (const __nv_bfloat16*)w2,
                        (const __nv_bfloat16*)b2, (__nv_bfloat16*)y_out, negative_slope, 0};
  return conv3x3_launch_chain(P, n, h, w, dtype, st, "conv3x3_ca_forward");
}

extern "C" int eavsr_conv3x3_chain_forward(const EavsrConvLayer* layers, int nlayers, int n, int h, int w, int dtype,
                                           void* sync_workspace, void* stream) {
  EAVSR_REQUIRE(layers && nlayers >= 1 && nlayers <= CV_MAXL, "conv3x3_chain_forward: 1..%d layers (got %d)", CV_MAXL, nlayers);
  static thread_local ChainParams P;
  P.nlayers = nlayers;
  P.sync = (unsigned*)sync_workspace;
  for (int i = 0; i < nlayers; ++i) {
    const EavsrConvLayer& a = layers[i];
    P.L[i] = ConvLayerDev{(const __nv_bfloat16*)a.x, (const uint8_t*)a.packed_weight, (const __nv_bfloat16*)a.bias,
                          (__nv_bfloat16*)a.out, a.channel_sums, (const __nv_bfloat16*)a.res, a.res_sums,
                          (const __nv_bfloat16*)a.w1, (const __nv_bfloat16*)a.b1, (const __nv_bfloat16*)a.w2,
                          (const __nv_bfloat16*)a.b2, (__nv_bfloat16*)a.y_out, a.negative_slope, 0};
  }
  return conv3x3_launch_chain(P, n, h, w, dtype, (cudaStream_t)stream, "conv3x3_chain_forward");
}

#ifdef EAVSR_CONV_TRACE
extern "C" int eavsr_debug_conv_trace(unsigned long long* host, int count) {
  return (int)cudaMemcpyFromSymbol(host, g_conv_trace, sizeof(unsigned long long) * count);
}
extern "C" int eavsr_debug_conv_trace2(unsigned long long* host, int count) {
  return (int)cudaMemcpyFromSymbol(host, g_conv_trace2, sizeof(unsigned long long) * count);
}
#endif
