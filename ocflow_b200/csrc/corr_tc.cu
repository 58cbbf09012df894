// Local cost volume (d = 4, 81 displacements) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the
// accumulator in tensor memory, fp32 parity kept by the 3xTF32 operand split.
//
//   out[b, k(dy,dx), y, x] = 1/C * sum_c f1[b,c,y,x] * f2[b,c,y+dy,x+dx]        (correlation_layer.py:7-40)
//
// GEMM view of one CTA tile: M = 128 pixels (16 rows x 8 columns), N = 384 halo pixels (24 x 16), K = channels.
// D = F1^T F2 contains, for every pixel, the products with ALL 384 halo pixels; the 81 wanted ones form a band that the
// epilogue cuts out.  The band wastes 384/81 = 4.7x of the tensor-core work and the operand split another 3x: 14x the useful
// flops, i.e. 1.1 PFLOP/s of dense TF32 is worth ~78 TFLOP/s here -- the fp32 FMA peak.  Measured, see below and DESIGN.md.
//
// 3xTF32: x = hi + lo with hi = tf32(x) (round to nearest, low 13 mantissa bits zero) and lo = tf32(x - hi);
// a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi, fp32 accumulation in TMEM.  Dropped terms are ~2^-22 |a||b| per product
// (measured against the fp64 oracle: 1e-6 relative, the FMA kernel has 1e-7; the bar is 1e-4).
//
// Roles in the 320-thread CTA (one CTA per SM, persistent over tiles):
//   warp 9     TMA issuer  (16-byte aligned rows) one thread: raw fp32 boxes [channel][row][x] of the f1 tile and the f2 halo, zero filled
//                          outside the image / past the last channel by the TMA unit, into a 4-deep raw ring
//   warps 4-7  converters  raw stage -> registers (conflict-free LDS.128) -> optional normalisation (x - mean) * inv_std (zero
//                          padding stays zero AFTER it, as normalize_features + F.pad give, correlation_layer.py:42-82) -> tf32
//                          hi / lo split -> 4x4 register transpose -> operand stage in the UMMA K-major no-swizzle canonical
//                          layout (16-byte chunks = 4 channels of one pixel).  Ragged rows (W % 4 != 0), which TMA cannot
//                          describe, use the same warps as software producers (predicated LDG, two stages of register prefetch)
//   warp 8     MMA issuer  one thread: 6 x tcgen05.mma (M128 N192 K8) per 8-channel stage, tcgen05.commit -> mbarriers
//   warps 0-3  epilogue    tcgen05.ld (TMEM lane = pixel) -> band selection by ADDRESS (the register index of a column is
//                          static, the output plane it belongs to is lane dependent) -> staging tile in shared memory ->
//                          1/C, LeakyReLU, sign bitmask, 128-bit coalesced stores
//
// Where the time goes (B200, 8x128x96x128, per 8-channel stage of one tile; developer builds that disabled one role at a time):
// full kernel 1900 cycles; converters idle (stages handed straight on) 1330; one MMA instead of three 1710; no converter stores
// 1600; the tensor-core work itself is 576 cycles.  I.e. the kernel is bound by the L2 -> shared-memory feed of the raw boxes
// (32- and 64-byte box rows, 4x halo amplification: 4.7 TB/s of L2 reads), then by the converters; the fp32 FMA kernels of
// corr.cu need the same time at C = 128 and less at C <= 64, which is why they are the default where TMA can feed them.
#include <cuda.h>  // CUtensorMap (types only)
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "corr_tc.cuh"

namespace {

constexpr int D = 4, ND = 9, NP = 81;
constexpr int TH = 16, TW = 8;             // pixel tile (M = 128, m = y * 8 + x)
constexpr int HROWS = TH + 2 * D;          // 24 halo rows
constexpr int NHALF = HROWS * 8;           // 192 accumulator columns per x-half of the halo (n = half * 192 + row * 8 + x % 8)
constexpr int KC = 8;                      // channels per stage = K of one kind::tf32 MMA
constexpr int STAGES = 4;
// K-major no-swizzle canonical layout (pinned by tools/tc_probe.cu): element (mn, k) at (mn / 8) * SBO + (k / 4) * LBO + (mn % 8) * 16
// + (k % 4) * 4 bytes -- core matrices of 8 rows x 16 B.  SBO carries 16 B of padding and the second x-half of the halo another
// 32 B, so that the 8 lanes of a quarter-warp -- which cover 128 CONTIGUOUS bytes of a raw TMA box (2 halo rows x 4 chunks, or
// 4 tile rows x 2 chunks: conflict-free 128-bit loads) -- also store to 8 distinct 16-byte bank groups: the group of a store is
// (row + 4 * (chunk & 1) + 2 * (x-half) + pixel) mod 8.
constexpr int LBO = 128, SBO = 272;
constexpr int A_BYTES = TH * SBO;              // 16 row groups (one tile row of 8 pixels each): 4352
constexpr int BHALF_BYTES = HROWS * SBO + 32;  // one x-half of the halo: 24 row groups + pad = 6560
constexpr int B_BYTES = 2 * BHALF_BYTES;       // 13120
constexpr int OFF_ALO = A_BYTES, OFF_BHI = 2 * A_BYTES, OFF_BLO = 2 * A_BYTES + B_BYTES;
constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);  // 34944 (a multiple of 128)
static_assert(STAGE_BYTES % 128 == 0, "stages must keep the raw ring 128-byte aligned");
constexpr int PS = 132;                    // floats per staged output plane (128 pixels + 4: 16-byte aligned rows, <= 2-way bank conflicts)
constexpr int STAGING_BYTES = NP * PS * 4;
constexpr int THREADS = 320;               // warps 0-3 epilogue, 4-7 producers / converters, 8 MMA issuer, 9 TMA issuer (TMA-fed variant)
constexpr int TMEM_COLS = 512;
// TMA-fed variant (16-byte aligned rows): raw fp32 boxes [channel][row][x] of the f1 tile (8 x 16 x 8) and the f2 halo
// (8 x 24 x 16) land in a ring of their own; the converter warps turn a raw stage into an operand stage (hi / lo, K-major)
constexpr int RAW_A_BYTES = KC * TH * TW * 4;            // 4096
constexpr int RAW_B_BYTES = KC * HROWS * 16 * 4;         // 12288
constexpr int RAW_BYTES = RAW_A_BYTES + RAW_B_BYTES;     // 16384
constexpr int RSTAGES = 4;
constexpr int STAGES_TMA = 3;                            // operand stages of the TMA-fed variant (shared memory budget)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned ok;
#ifdef OCF_TC_WATCHDOG
  unsigned spins = 0;
#endif
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
#ifdef OCF_TC_WATCHDOG
    if (!ok && ++spins > (1u << 22)) __trap();   // developer builds: turn a protocol bug into an error instead of a hang
#endif
  } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one 4-D box {x, y, c, b} of an NCHW fp32 tensor -> dense [c][y][x] box in shared memory (zero fill outside the tensor)
__device__ __forceinline__ void tma_load_4d(unsigned smem_dst, const CUtensorMap* map, unsigned long long* bar, int x, int y, int c, int b) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(c), "r"(b)
      : "memory");
}
__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: start address, leading byte offset (between the two 4-channel core
// matrices of one K = 8 step), stride byte offset (between groups of 8 rows)
__device__ __forceinline__ unsigned long long umma_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes) {
  return (unsigned long long)((addr >> 4) & 0x3FFF) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(unsigned d_tmem, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// round to the nearest tf32 (10 mantissa bits) with integer ops at full rate; the tensor core TRUNCATES fp32 operands to tf32
// (measured, tools/tc_probe.cu), so the rounding has to happen here for the hi / lo split to be exact
__device__ __forceinline__ float tf32_round(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void sts128(unsigned addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct TileCoord {
  int b, y0, x0;
};
__device__ __forceinline__ TileCoord tile_coord(int tile, int tiles_x, int tiles_y) {
  TileCoord t;
  const int txi = tile % tiles_x, r = tile / tiles_x;
  t.x0 = txi * TW;
  t.y0 = (r % tiles_y) * TH;
  t.b = r / tiles_y;
  return t;
}

// VEC: rows are 16-byte aligned (W % 4 == 0, aligned base pointers) -> one LDG.128 per 4-pixel chunk and 128-bit output stores;
// otherwise (KITTI / Sintel pyramid widths) per-element predicated loads and stores: no re-pitching copy is needed.
// TMAP: the TMA-fed variant (requires VEC).  The software producers are latency bound (one L2 round trip per stage and thread
// even with two stages of register prefetch: ncu long_scoreboard, 3150 cycles per stage against 576 cycles of tensor-core
// work); with TMA the global latency is covered by a 4-deep raw ring that costs no registers and no issue slots.
template <bool VEC, bool TMAP>
__global__ void __launch_bounds__(THREADS, 1)
corr_fwd_tc_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                   const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ out, unsigned char* __restrict__ mask,
                   const float* __restrict__ norm, float* __restrict__ f1n_out, long long f1n_bstride, float* __restrict__ f2n_out, int C, int H,
                   int W, long long out_bstride, float inv_c, float slope, int tiles_x, int tiles_y, int ntiles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full_bar[STAGES], empty_bar[STAGES], tmem_full_bar, tmem_empty_bar[2];
  __shared__ __align__(8) unsigned long long raw_full[RSTAGES], raw_empty[RSTAGES];
  __shared__ unsigned tmem_base_slot;
  constexpr int NSTG = TMAP ? STAGES_TMA : STAGES;          // operand stages in use
  // layout: operand ring | raw ring (TMAP only; TMA destinations must be 128-byte aligned) | output staging tile
  const unsigned raw_base = smem_u32(smem) + (unsigned)(NSTG * STAGE_BYTES);
  float* staging = reinterpret_cast<float*>(smem + NSTG * STAGE_BYTES + (TMAP ? RSTAGES * RAW_BYTES : 0));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nst = (C + KC - 1) / KC;
  const size_t HW = (size_t)H * W;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 1); }
#pragma unroll
    for (int s = 0; s < RSTAGES; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 128); }
    mbar_init(&tmem_full_bar, 1);
    mbar_init(&tmem_empty_bar[0], 128);
    mbar_init(&tmem_empty_bar[1], 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tb = tmem_base_slot;

  if (warp >= 4 && warp < 8) {
    // =========================== producers ===========================
    // A unit of work = 4 channels x 4 adjacent pixels (4 x LDG.128 along x), transposed in registers into 4 x STS.128 of
    // [4 channels] per pixel: the K-major canonical layout wants the channels (K) of one pixel contiguous.  256 units per stage
    // (64 for the f1 tile, 192 for the f2 halo), two per thread.  Lane bits 0-2 = (channel group, chunk parity, row parity):
    // with LBO = 144 B and SBO = 288 B these 8 lanes hit 8 distinct 16-byte bank groups (conflict-free STS.128).
    const int p = tid - 128;
    // unit 0: f2 unit p ; unit 1: f2 unit 128 + p (p < 64) or f1 unit p - 64.  f2 units: ub = kg * 96 + row * 4 + chunk (24 halo
    // rows x 4 chunks of 4 pixels), f1 units: ua = kg * 32 + row * 2 + chunk (16 tile rows x 2 chunks).
    int kg[2], row[2], jx[2];
    bool isa[2];
    {
      isa[0] = false; isa[1] = p >= 64;
      const int ub0 = p;
      kg[0] = ub0 >= 96 ? 1 : 0;
      { const int r = ub0 - kg[0] * 96; row[0] = r >> 2; jx[0] = r & 3; }
      if (isa[1]) { const int ua = p - 64; kg[1] = ua >> 5; row[1] = (ua & 31) >> 1; jx[1] = ua & 1; }
      else { const int r = 128 + p - 96; kg[1] = 1; row[1] = r >> 2; jx[1] = r & 3; }
    }
    // shared-memory offset of pixel 0 of the unit (hi plane): row-group * SBO + channel-group * LBO + (x % 8) * 16
    unsigned soff[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (isa[i]) soff[i] = (unsigned)(row[i] * SBO + kg[i] * LBO + (4 * jx[i]) * 16);
      else soff[i] = (unsigned)(OFF_BHI + (jx[i] >> 1) * BHALF_BYTES + row[i] * SBO + kg[i] * LBO + (4 * (jx[i] & 1)) * 16);
    }
    float nmean = 0.f, ninv = 1.f;
    if (norm != nullptr) { nmean = __ldg(norm); ninv = __ldg(norm + 1); }
    const int ntl = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = ntl * nst;

    // global loads of stage n into registers (no dependence on the ring: they run PF stages ahead of the stores)
    auto load_stage = [&](int n, float4 (&v)[8], unsigned& okm) {
      const int it = n / nst, s = n - it * nst;
      const TileCoord tc = tile_coord((int)blockIdx.x + it * (int)gridDim.x, tiles_x, tiles_y);
      okm = 0u;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int y = isa[i] ? tc.y0 + row[i] : tc.y0 - D + row[i];
        const int x = isa[i] ? tc.x0 + 4 * jx[i] : tc.x0 - D + 4 * jx[i];
        const float* base = (isa[i] ? f1 : f2) + (size_t)tc.b * C * HW;
        const bool rok = y >= 0 && y < H;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = s * KC + kg[i] * 4 + q;
          const float* src = base + ((size_t)c * H + y) * W + x;
          float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
          unsigned m = 0u;
          if (rok && c < C) {
            if (VEC) {
              if (x >= 0 && x < W) { r = __ldg(reinterpret_cast<const float4*>(src)); m = 0xFu; }
            } else {
              if (x >= 0 && x < W) { r.x = __ldg(src); m |= 1u; }
              if (x + 1 >= 0 && x + 1 < W) { r.y = __ldg(src + 1); m |= 2u; }
              if (x + 2 >= 0 && x + 2 < W) { r.z = __ldg(src + 2); m |= 4u; }
              if (x + 3 >= 0 && x + 3 < W) { r.w = __ldg(src + 3); m |= 8u; }
            }
          }
          v[i * 4 + q] = r;
          okm |= m << ((i * 4 + q) * 4);
        }
      }
    };
    // normalise, split into tf32 hi / lo, transpose 4 channels x 4 pixels and store stage n
    auto store_stage = [&](int n, const float4 (&v)[8], unsigned okm) {
      const int slot = n % NSTG;
      if (n >= NSTG) mbar_wait(&empty_bar[slot], ((n / NSTG) - 1) & 1);
      const unsigned sbase = smem_u32(smem) + (unsigned)slot * STAGE_BYTES;
      const int it = n / nst, s = n - it * nst;
      const TileCoord tc = tile_coord((int)blockIdx.x + it * (int)gridDim.x, tiles_x, tiles_y);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float e[4][4];   // [channel q][pixel]
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 r = v[i * 4 + q];
          e[q][0] = r.x; e[q][1] = r.y; e[q][2] = r.z; e[q][3] = r.w;
        }
        if (norm != nullptr) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const unsigned m = (okm >> ((i * 4 + q) * 4)) & 0xFu;
#pragma unroll
            for (int px = 0; px < 4; ++px) e[q][px] = (m >> px) & 1u ? (e[q][px] - nmean) * ninv : 0.f;
            // the normalised feature maps are outputs of the level: c1n is concatenated into the decoder input
            // (cost_volume_flow_net.py:190), c2n is kept for the backward.  Every f1 element and the interior of the f2 halo
            // box (rows 4..19, chunks 1 and 2 = this tile's own 16 x 8 pixels) are loaded by exactly one tile.
            float* dst = nullptr;
            const int c = s * KC + kg[i] * 4 + q;
            if (isa[i]) {
              if (f1n_out != nullptr)
                dst = f1n_out + (size_t)tc.b * (f1n_bstride ? (size_t)f1n_bstride : (size_t)C * HW) + ((size_t)c * H + tc.y0 + row[i]) * W + tc.x0 + 4 * jx[i];
            } else if (f2n_out != nullptr && (jx[i] == 1 || jx[i] == 2) && row[i] >= D && row[i] < D + TH) {
              dst = f2n_out + (size_t)tc.b * C * HW + ((size_t)c * H + tc.y0 - D + row[i]) * W + tc.x0 - D + 4 * jx[i];
            }
            if (dst != nullptr && m) {
              if (VEC) *reinterpret_cast<float4*>(dst) = make_float4(e[q][0], e[q][1], e[q][2], e[q][3]);
              else {
#pragma unroll
                for (int px = 0; px < 4; ++px) if ((m >> px) & 1u) dst[px] = e[q][px];
              }
            }
          }
        }
        const unsigned a = sbase + soff[i];
        const unsigned lo_off = isa[i] ? (unsigned)A_BYTES : (unsigned)B_BYTES;
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          float hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) { hi[q] = tf32_round(e[q][px]); lo[q] = tf32_round(e[q][px] - hi[q]); }
          sts128(a + px * 16, hi[0], hi[1], hi[2], hi[3]);
          sts128(a + px * 16 + lo_off, lo[0], lo[1], lo[2], lo[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(&full_bar[slot]);
    };

    if constexpr (TMAP) {
      // converters: raw stage (TMA, [channel][row][x]) -> registers -> operand stage.  Validity (for the normalisation: padding
      // stays zero) comes from the coordinates; the TMA unit has already zero-filled everything outside the tensor.
      unsigned roff[2];   // byte offset of channel 0 of the unit's 4-channel group inside a raw stage
      unsigned cstep[2];  // bytes between channels
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (isa[i]) { cstep[i] = TH * TW * 4; roff[i] = (unsigned)(((kg[i] * 4) * TH + row[i]) * TW + 4 * jx[i]) * 4u; }
        else { cstep[i] = HROWS * 16 * 4; roff[i] = (unsigned)RAW_A_BYTES + (unsigned)(((kg[i] * 4) * HROWS + row[i]) * 16 + 4 * jx[i]) * 4u; }
      }
      for (int n = 0; n < total; ++n) {
        const int rs = n % RSTAGES;
        mbar_wait(&raw_full[rs], (n / RSTAGES) & 1);
        const unsigned rb = raw_base + (unsigned)rs * RAW_BYTES;
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) v[i * 4 + q] = lds128(rb + roff[i] + (unsigned)q * cstep[i]);
        unsigned okm = 0xFFFFFFFFu;
        if (norm != nullptr) {
          const int it = n / nst, s = n - it * nst;
          const TileCoord tc = tile_coord((int)blockIdx.x + it * (int)gridDim.x, tiles_x, tiles_y);
          okm = 0u;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int y = isa[i] ? tc.y0 + row[i] : tc.y0 - D + row[i];
            const int x = isa[i] ? tc.x0 + 4 * jx[i] : tc.x0 - D + 4 * jx[i];
            const bool ok = y >= 0 && y < H && x >= 0 && x < W;     // W % 4 == 0: a 4-pixel chunk is entirely inside or outside
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (ok && s * KC + kg[i] * 4 + q < C) okm |= 0xFu << ((i * 4 + q) * 4);
          }
        }
        store_stage(n, v, okm);
        mbar_arrive(&raw_empty[rs]);   // after the values have been consumed: the TMA unit may refill the raw stage
      }
    } else {
    // software pipeline: the loads of stage n + 2 are in flight while stage n is converted and stored (3 register buffers)
    float4 v0[8], v1[8], v2[8];
    unsigned m0 = 0u, m1 = 0u, m2 = 0u;
    if (total > 0) load_stage(0, v0, m0);
    if (total > 1) load_stage(1, v1, m1);
    for (int n = 0; n < total; n += 3) {
      if (n + 2 < total) load_stage(n + 2, v2, m2);
      store_stage(n, v0, m0);
      if (n + 1 < total) {
        if (n + 3 < total) load_stage(n + 3, v0, m0);
        store_stage(n + 1, v1, m1);
      }
      if (n + 2 < total) {
        if (n + 4 < total) load_stage(n + 4, v1, m1);
        store_stage(n + 2, v2, m2);
      }
    }
    }
  } else if (warp == 9) {
    // =========================== TMA issuer (TMA-fed variant) ===========================
    if (TMAP && lane == 0) {
      const int ntl = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      int n = 0;
      for (int it = 0; it < ntl; ++it) {
        const TileCoord tc = tile_coord((int)blockIdx.x + it * (int)gridDim.x, tiles_x, tiles_y);
        for (int s = 0; s < nst; ++s, ++n) {
          const int rs = n % RSTAGES;
          if (n >= RSTAGES) mbar_wait(&raw_empty[rs], ((n / RSTAGES) - 1) & 1);
          const unsigned rb = raw_base + (unsigned)rs * RAW_BYTES;
          mbar_expect_tx(&raw_full[rs], RAW_BYTES);
          tma_load_4d(rb, &map1, &raw_full[rs], tc.x0, tc.y0, s * KC, tc.b);
          tma_load_4d(rb + RAW_A_BYTES, &map2, &raw_full[rs], tc.x0 - D, tc.y0 - D, s * KC, tc.b);
        }
      }
    }
  } else if (warp == 8) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B tf32, both K-major, N = 192, M = 128
      constexpr unsigned IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(NHALF >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
      int n = 0, t = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
        for (int s = 0; s < nst; ++s, ++n) {
          const int slot = n % NSTG;
          mbar_wait(&full_bar[slot], (n / NSTG) & 1);
          tc_fence_after();
          const unsigned sbase = smem_u32(smem) + (unsigned)slot * STAGE_BYTES;
          const unsigned long long ahi = umma_desc(sbase, LBO, SBO), alo = umma_desc(sbase + OFF_ALO, LBO, SBO);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (s == 0) {   // the epilogue must have drained this half of the previous tile's accumulator
              mbar_wait(&tmem_empty_bar[h], (t & 1) ^ 1);
              tc_fence_after();
            }
            const unsigned long long bhi = umma_desc(sbase + OFF_BHI + h * BHALF_BYTES, LBO, SBO);
            const unsigned long long blo = umma_desc(sbase + OFF_BLO + h * BHALF_BYTES, LBO, SBO);
            const unsigned d = tb + (unsigned)(h * NHALF);
            umma_tf32(d, ahi, bhi, IDESC, s > 0 ? 1u : 0u);
            umma_tf32(d, ahi, blo, IDESC, 1u);
            umma_tf32(d, alo, bhi, IDESC, 1u);
          }
          tc_commit(&empty_bar[slot]);   // arrives when the MMAs above have finished reading the stage
        }
        tc_commit(&tmem_full_bar);       // ... and when the whole tile is accumulated
      }
    }
  } else {
    // =========================== epilogue (warps 0-3: TMEM lanes 32 w .. 32 w + 31) ===========================
    const int pyl = lane >> 3, px = lane & 7;
    const int m = warp * 32 + lane;                       // pixel within the tile == TMEM lane
    const size_t bstride = out_bstride ? (size_t)out_bstride : (size_t)NP * HW;
    const int Wb = (W + 7) >> 3;
    int t = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
      const TileCoord tc = tile_coord(tile, tiles_x, tiles_y);
      mbar_wait(&tmem_full_bar, t & 1);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int bt = 0; bt < 3; ++bt) {   // 4 halo rows per load: rows 4 w + 4 bt .. + 3 (the warp's pixel rows need 4 w .. 4 w + 11)
          unsigned r[32];
          tmem_ld32(tb + ((unsigned)(warp * 32) << 16) + (unsigned)(h * NHALF + (warp * 4 + bt * 4) * 8), r);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int dy = bt * 4 + q - pyl;
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              const int dx = 8 * h + cc - px;
              if (dy >= 0 && dy < ND && dx >= 0 && dx < ND) staging[(dy * ND + dx) * PS + m] = __uint_as_float(r[q * 8 + cc]);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[h]);   // this half of the accumulator may be overwritten by the next tile
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // coalesced copy-out: 81 planes x 32 float4 (4 pixels of one tile row each)
      for (int idx = tid; idx < NP * 32; idx += 128) {
        const int plane = idx >> 5, q4 = idx & 31;
        const int row = q4 >> 1, xq = (q4 & 1) << 2;
        const float4 v = *reinterpret_cast<const float4*>(staging + plane * PS + q4 * 4);
        float e[4] = {v.x * inv_c, v.y * inv_c, v.z * inv_c, v.w * inv_c};
        unsigned nib = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          nib |= (e[j] > 0.f ? 1u : 0u) << j;
          e[j] = e[j] > 0.f ? e[j] : e[j] * slope;
        }
        const unsigned other = __shfl_xor_sync(0xffffffffu, nib, 1);   // idx and idx ^ 1 are the two halves of one 8-pixel row
        const int y = tc.y0 + row, x = tc.x0 + xq;
        if (y < H && x < W) {
          if (mask != nullptr && xq == 0) mask[(((size_t)tc.b * NP + plane) * H + y) * Wb + (x >> 3)] = (unsigned char)(nib | (other << 4));
          float* o = out + (size_t)tc.b * bstride + ((size_t)plane * H + y) * W + x;
          if (VEC) *reinterpret_cast<float4*>(o) = make_float4(e[0], e[1], e[2], e[3]);
          else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (x + j < W) o[j] = e[j];
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // staging tile free for the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(TMEM_COLS));
}

}  // namespace

int ocf_corr_fwd_tc_launch(const float* f1, const float* f2, float* out, unsigned char* mask_out, const float* norm, float* f1n_out,
                           long long f1n_bstride, float* f2n_out, int B, int C, int H, int W, long long out_bstride, float leaky_slope,
                           cudaStream_t s) {
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const long long ntiles = (long long)tiles_x * tiles_y * B;
  if (ntiles > 0x7fffffffLL) return OCF_EUNSUPPORTED;
  const bool vec = (W % 4 == 0) && ocf_aligned16(f1) && ocf_aligned16(f2) && ocf_aligned16(out) && (out_bstride % 4 == 0) &&
                   (f1n_out == nullptr || (ocf_aligned16(f1n_out) && f1n_bstride % 4 == 0)) && (f2n_out == nullptr || ocf_aligned16(f2n_out));
  static const int no_tma = []() { const char* e = getenv("OCF_TC_NO_TMA"); return e ? atoi(e) : 0; }();   // developer knob
  CUtensorMap m1, m2;
  memset(&m1, 0, sizeof(m1));
  memset(&m2, 0, sizeof(m2));
  const bool tma = vec && !no_tma && ocf_make_tensor_map(&m1, f1, B, C, H, W, TW, TH, KC, 0) && ocf_make_tensor_map(&m2, f2, B, C, H, W, 16, HROWS, KC, 0);
  const size_t smem = tma ? (size_t)STAGES_TMA * STAGE_BYTES + STAGING_BYTES + (size_t)RSTAGES * RAW_BYTES : (size_t)STAGES * STAGE_BYTES + STAGING_BYTES;
  auto kernel = tma ? corr_fwd_tc_kernel<true, true> : (vec ? corr_fwd_tc_kernel<true, false> : corr_fwd_tc_kernel<false, false>);
  static bool attr_set[3] = {false, false, false};
  const int ki = tma ? 2 : (vec ? 1 : 0);
  if (!attr_set[ki]) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set[ki] = true;
  }
  const int grid = ntiles < OCF_SM_COUNT ? (int)ntiles : OCF_SM_COUNT;
  kernel<<<grid, THREADS, smem, s>>>(m1, m2, f1, f2, out, mask_out, norm, f1n_out, f1n_bstride, f2n_out, C, H, W, out_bstride, 1.0f / (float)C, leaky_slope,
                                     tiles_x, tiles_y, (int)ntiles);
  return ocf_launch_status();
}
