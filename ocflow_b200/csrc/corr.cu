// Local cost volume (PWC-style correlation), forward and backward, fp32, NCHW.
//
// Replaces compute_cost_volume (reference models/networks/correlation_layer.py:7-40): 81 x
// (slice, mul, mean) + cat => one launch.  See DESIGN.md "corr" for the tiling rationale.
//
//   staging : when rows are 16-byte aligned (W % 4 == 0: FlyingChairs / any multiple-of-64 crop) the f1 tile and the
//             f2 tile + halo of a channel chunk are ONE TMA box each (cp.async.bulk.tensor.4d over the NCHW tensor,
//             out-of-image and past-the-last-channel elements zero-filled by the TMA unit), issued by a single
//             producer thread into a STAGES-deep ring guarded by full/empty mbarriers -- the 9 compute warps never
//             execute a staging instruction or a __syncthreads in the main loop.  Ragged widths (KITTI/Sintel
//             pyramids: 621, 311, 109 ... where TMA's 16-byte global strides are illegal) use a cp.async ring with
//             element-wise zero-fill instead.
//   forward : a CTA owns a TH x TW pixel tile of one batch item (and, for the small pyramid levels, one
//             slice of the channels).  Shared-memory row strides are == 4 mod 8 floats so the 128-bit window
//             loads of a quarter-warp hit 8 distinct bank groups.  Warp w owns the vertical displacement dy = w - d; a thread owns PX consecutive pixels
//             of one row and all 2d+1 horizontal displacements: (2d+1)*PX register accumulators,
//             (PX + PX+2d)/4 LDS.128 per (2d+1)*PX FMAs.
//             Small levels (few tiles, many channels) split the channels over a thread-block CLUSTER; the
//             partial cost volumes are reduced through distributed shared memory (deterministic, no atomics,
//             the 1/C scale and the LeakyReLU stay fused).
//   backward: d f1 and d f2 are both written as gathers (deterministic, no atomics):
//               d f1[c,p] = 1/C sum_delta g[delta, p]        * f2[c, p+delta]
//               d f2[c,q] = 1/C sum_delta g[-delta, q+delta] * f1[c, q+delta]
//             i.e. the same "window of the other feature times 81 per-pixel coefficients"; only the
//             way the coefficients are fetched differs.  A thread keeps its (2d+1)*PX coefficients
//             in registers for the whole kernel, the 2d+1 dy-warps reduce through shared memory.  Channels
//             are independent here, so small levels simply split them over more CTAs.
#include <cooperative_groups.h>
#include <cuda.h>  // CUtensorMap types only; the driver entry point is fetched at run time (no libcuda link dependency)

#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "corr_tc.cuh"

namespace cg = cooperative_groups;

// persistent forward kernel: registers per thread, channel unroll and resident CTAs per SM (tuning knobs)
#ifndef OCF_FWD_REGS
#define OCF_FWD_REGS 168
#endif
#ifndef OCF_FWD_UNROLL
#define OCF_FWD_UNROLL 8
#endif
#ifndef OCF_FWD_CTAS_PER_SM
#define OCF_FWD_CTAS_PER_SM 1
#endif
// tiled backward, TMA path: specialised cross-dy reduction (32-bit indices, packed adds); 0 = the general loop (A/B builds)
#ifndef OCF_BWD_FASTRED
#define OCF_BWD_FASTRED 1
#endif
// tiled backward, TMA path: one mbarrier per coefficient box (a warp lifts as soon as its own box has landed); 0 = one for all
#ifndef OCF_BWD_GBAR_PER_BOX
#define OCF_BWD_GBAR_PER_BOX 1
#endif

#ifdef OCF_TIMELINE
// developer builds only (tools/timeline.py): per-CTA globaltimer stamps of the persistent kernels
__device__ unsigned long long ocf_tl[1024 * 16];
__device__ __forceinline__ unsigned long long ocf_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define OCF_TL(slot) do { if (threadIdx.x == 0 && blockIdx.x < 1024) ocf_tl[blockIdx.x * 16 + (slot)] = ocf_now(); } while (0)
#define OCF_TLX(cond, slot) do { if ((cond) && blockIdx.x < 1024) ocf_tl[blockIdx.x * 16 + (slot)] = ocf_now(); } while (0)
#define OCF_TL_SM() do { if (threadIdx.x == 0 && blockIdx.x < 1024) { unsigned sm; asm("mov.u32 %0, %smid;" : "=r"(sm)); ocf_tl[blockIdx.x * 16 + 15] = sm; } } while (0)
// 3-D grids (tiled backward): rows indexed by the linear block id
#define OCF_TLB(cond, slot) do { const unsigned l_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); \
    if ((cond) && l_ < 1024) ocf_tl[l_ * 16 + (slot)] = ocf_now(); } while (0)
#define OCF_TLB_VAL(cond, slot, val) do { const unsigned l_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); \
    if ((cond) && l_ < 1024) ocf_tl[l_ * 16 + (slot)] = (unsigned long long)(val); } while (0)
extern "C" int ocf_debug_timeline(unsigned long long* host, int n) {
  if (host == nullptr) {   // reset (between two settings probed in one process)
    void* p = nullptr;
    cudaError_t e = cudaGetSymbolAddress(&p, ocf_tl);
    return e != cudaSuccess ? (int)e : (int)cudaMemset(p, 0, sizeof(ocf_tl));
  }
  return (int)cudaMemcpyFromSymbol(host, ocf_tl, sizeof(unsigned long long) * n);
}
#else
#define OCF_TL(slot) do { } while (0)
#define OCF_TLX(cond, slot) do { } while (0)
#define OCF_TL_SM() do { } while (0)
#define OCF_TLB(cond, slot) do { } while (0)
#define OCF_TLB_VAL(cond, slot, val) do { } while (0)
#endif

namespace {

constexpr int pad4mod8(int n) {
  int r = (n + 3) / 4 * 4;
  return (r % 8 == 4) ? r : r + 4;
}

// NDY: vertical displacements handled by one CTA (one warp each).  d = 4: all 9.  d = 10 (FlowNetC family, 441 planes): 7 of the
// 21, i.e. three CTAs per tile, each staging only the TH + 6 halo rows its displacements touch -- 21 warps x 84 accumulators would
// not fit the register file.  MAXREG: register cap of the forward kernel (threads x registers must allow 2 CTAs per SM).
template <int D_, int PX_, int TXT_, int TH_, int CC_, int STAGES_, int NDY_ = 2 * D_ + 1, int MAXREG_ = 96>
struct CorrTile {
  static constexpr int D = D_, PX = PX_, TXT = TXT_, TH = TH_, CC = CC_, STAGES = STAGES_, NDY = NDY_, MAXREG = MAXREG_;
  static constexpr int ND = 2 * D + 1;
  static constexpr int NGY = ND / NDY;          // CTAs (dy groups) per tile
  static_assert(ND % NDY == 0, "the dy groups must tile the displacement range");
  static constexpr int TW = PX * TXT;
  static constexpr int F2W = TW + 2 * D;
  static constexpr int F2H = TH + NDY - 1;      // halo rows of one dy group (== TH + 2 D when NDY == ND)
  static constexpr int S1 = pad4mod8(TW);
  static constexpr int S2 = pad4mod8(F2W);
  static constexpr int LANES = TXT * TH;  // pixel-threads per dy (one warp when == 32)
  static constexpr int THREADS = LANES * NDY;
  static constexpr int WIN = PX + 2 * D;  // f2 window per thread
  static constexpr int F1_STAGE = CC * TH * S1;    // floats per stage
  static constexpr int F2_STAGE = CC * F2H * S2;
  static_assert(PX % 4 == 0 && (2 * D) % 4 == 0, "128-bit window loads need PX and 2D multiples of 4");
  static_assert(LANES == 32, "one warp per vertical displacement");
};

// ---- cp.async helpers ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;  // src-size 0 => the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMA + mbarrier helpers ---------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
// one 4-D box {x, y, c, b} of an NCHW fp32 tensor -> dense [c][y][x] box in shared memory; signals `bar` with the byte count
__device__ __forceinline__ void tma_load_4d(float* smem_dst, const CUtensorMap* map, unsigned long long* bar, int x, int y, int c,
                                            int b) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(c), "r"(b)
      : "memory");
}
// the same box, only as far as the L2 (a hint: no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int x, int y, int c, int b) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(c), "r"(b)
               : "memory");
}
// dense [c][y][x] box in shared memory -> one 4-D box {x, y, c, b} of an NCHW fp32 tensor (elements outside the tensor are dropped)
__device__ __forceinline__ void tma_store_4d(const float* smem_src, const CUtensorMap* map, int x, int y, int c, int b) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(smem_src)), "r"(x), "r"(y), "r"(c), "r"(b)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- packed fp32 FMA (Blackwell FFMA2): two IEEE fp32 fused multiply-adds per issue slot ------------
typedef unsigned long long u64;
template <int V>
struct IntC {
  static constexpr int value = V;
};
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(u64& d, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ void fadd2(u64& d, u64 a) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(d) : "l"(a)); }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

template <int THREADS>
__device__ __forceinline__ void consumer_bar_sync() {  // named barrier 1: the compute warps only (the producer warp is not part)
  asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
}

__device__ __forceinline__ void cp_async8(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

// Asynchronously copies a [CC][ROWS][COLS] box of one batch item of a NCHW tensor (origin (yb, xb), may be negative
// or past the edge; channels [c0, c0+CC) clipped to c_end) into shared memory with row stride S; everything outside
// is zero-filled by the copy engine.  VW = floats per copy: 4 needs 16-byte aligned rows and box origin (W % 4 == 0, xb % 4 == 0),
// 2 needs them 8-byte aligned (W % 2 == 0, xb % 2 == 0: the d = 10 halo, which starts at x0 - 10), 1 nothing.
template <int CC, int ROWS, int COLS, int S, int THREADS, int VW>
__device__ __forceinline__ void stage_box_async(float* __restrict__ dst, const float* __restrict__ src_b, int c0, int c_end,
                                                int H, int W, int yb, int xb) {
  static_assert(VW == 1 || VW == 2 || VW == 4, "copy width");
  constexpr int CV = COLS / VW;
  constexpr int ITEMS = CC * ROWS * CV;
  static_assert(COLS % VW == 0, "box width must be a multiple of the copy width");
#pragma unroll 4
  for (int i = threadIdx.x; i < ITEMS; i += THREADS) {
    const int row = i / CV, q = i - row * CV;
    const int c = row / ROWS, r = row - c * ROWS;
    const int gy = yb + r, gx = xb + q * VW, gc = c0 + c;
    const bool ok = gc < c_end && gy >= 0 && gy < H && gx >= 0 && gx < W;   // a VW-aligned group is entirely inside or outside the row
    const float* src = ok ? src_b + ((size_t)gc * H + gy) * W + gx : src_b;
    float* d = dst + (c * ROWS + r) * S + q * VW;
    if (VW == 4) cp_async16(d, src, ok);
    else if (VW == 2) cp_async8(d, src, ok);
    else cp_async4(d, src, ok);
  }
}

// ---- staging modes ------------------------------------------------------------------------------
constexpr int STG_ASYNC4 = 0;   // cp.async, 4-byte elements (ragged widths)
constexpr int STG_ASYNC16 = 1;  // cp.async, 16-byte elements
constexpr int STG_TMA = 2;      // TMA boxes + mbarrier ring, dedicated producer warp

template <class T, int STG>
struct Threads {
  static constexpr int value = T::THREADS + (STG == STG_TMA ? 32 : 0);
};

struct ChunkRange {
  int begin, count;
};
// channel chunks [begin, begin+count) of this CTA: whole chunks only, so a TMA box never straddles two slices
__device__ __forceinline__ ChunkRange chunk_range(int C, int CC, int ksplit, int ks) {
  const int total = (C + CC - 1) / CC;
  const int per = (total + ksplit - 1) / ksplit;
  ChunkRange r;
  r.begin = ks * per;
  r.count = max(0, min(total, r.begin + per) - r.begin);
  return r;
}

// ---- forward ------------------------------------------------------------------------------------
// grid (tiles_x, tiles_y, B * ksplit); when ksplit > 1 the launch carries cluster dims (1, 1, ksplit).
template <class T, int STG>
__global__ void __maxnreg__(T::MAXREG)
corr_fwd_tiled(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2, const float* __restrict__ f1,
               const float* __restrict__ f2, float* __restrict__ out, unsigned char* __restrict__ mask, int C, int H, int W,
               long long out_bstride, float inv_c, float slope, int ksplit) {
  constexpr int D = T::D, PX = T::PX, ND = T::ND, TH = T::TH, TW = T::TW, CC = T::CC, STAGES = T::STAGES;
  constexpr int S1 = T::S1, S2 = T::S2, F2H = T::F2H, F2W = T::F2W, WIN = T::WIN;
  constexpr int STAGE = T::F1_STAGE + T::F2_STAGE;
  constexpr bool TMA = STG == STG_TMA;
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) unsigned long long full_bar[STAGES], empty_bar[STAGES];

  const int tid = threadIdx.x;
  const int lane = tid % T::LANES;
  const int tx = lane % T::TXT, ty = lane / T::TXT;
  const int dyi = tid / T::LANES;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  // blockIdx.z = (b * NGY + dy group) * ksplit + channel slice
  const int zz = blockIdx.z / ksplit, ks = blockIdx.z - zz * ksplit;
  const int b = zz / T::NGY, dy0 = (zz - b * T::NGY) * T::NDY;
  const ChunkRange cr = chunk_range(C, CC, ksplit, ks);
  const int nchunks = cr.count;

  // Accumulators.  out[dx][p] += a[p] * w[p + dx]: for a fixed pixel p the 2d+1 products share a[p] and walk consecutive
  // window elements, so they are issued as packed FFMA2 (scalar a[p] broadcast x aligned pair of w): pixel p pairs the
  // displacements (dx, dx+1) with p + dx even -- dx = 0,2,.. for even p, dx = 1,3,.. for odd p -- and keeps the one
  // left-over displacement (dx = 2d for even p, dx = 0 for odd p) in a scalar FFMA.  Per channel: D*PX FFMA2 + PX FFMA
  // instead of (2D+1)*PX FFMA; same IEEE results, same summation order.
  static_assert(PX % 2 == 0, "pixel pairs");
  u64 accp[PX][D];
  float accs[PX];
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    accs[p] = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) accp[p][j] = 0ull;
  }

  auto compute = [&](const float* st) {
    const float* p1 = st + ty * S1 + tx * PX;
    const float* p2 = st + T::F1_STAGE + (ty + dyi) * S2 + tx * PX;
#pragma unroll 1
    for (int c = 0; c < CC; ++c) {
      float a[PX];
      u64 w2[WIN / 2];
#pragma unroll
      for (int q = 0; q < PX / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(p1 + c * TH * S1 + 4 * q);
        a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int q = 0; q < WIN / 4; ++q) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p2 + c * F2H * S2 + 4 * q);
        w2[2 * q] = v.x; w2[2 * q + 1] = v.y;
      }
#pragma unroll
      for (int p = 0; p < PX; ++p) {
        const u64 ap = pack2(a[p], a[p]);
        const int first = p & 1;  // first paired displacement
#pragma unroll
        for (int j = 0; j < D; ++j) ffma2(accp[p][j], ap, w2[(p + first + 2 * j) / 2]);
        float lo, hi;
        if (first == 0) {  // left-over dx = 2D: w[p + 2D]
          unpack2(w2[(p + 2 * D) / 2], lo, hi);
          accs[p] = fmaf(a[p], lo, accs[p]);
        } else {           // left-over dx = 0: w[p]
          unpack2(w2[p / 2], lo, hi);
          accs[p] = fmaf(a[p], hi, accs[p]);
        }
      }
    }
  };
  // unpacked view for the epilogue
  auto acc_get = [&](int dx, int p) -> float {
    const int first = p & 1;
    if (first == 0 && dx == 2 * D) return accs[p];
    if (first == 1 && dx == 0) return accs[p];
    float lo, hi;
    unpack2(accp[p][(dx - first) / 2], lo, hi);
    return ((dx - first) & 1) ? hi : lo;
  };

  if constexpr (TMA) {
    // thread 0 doubles as the TMA producer (no extra warp: 288 threads leave 112 registers per thread at 2 CTAs/SM)
    constexpr unsigned BYTES = sizeof(float) * STAGE;
    auto issue_tma = [&](int i) {
      const int s = i % STAGES;
      float* st = smem + s * STAGE;
      const int c0 = (cr.begin + i) * CC;
      mbar_expect_tx(&full_bar[s], BYTES);
      tma_load_4d(st, &map1, &full_bar[s], x0, y0, c0, b);
      tma_load_4d(st + T::F1_STAGE, &map2, &full_bar[s], x0 - D, y0 - D + dy0, c0, b);
    };
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], T::NDY); }
      mbar_fence_init();
#pragma unroll
      for (int s = 0; s < STAGES; ++s)
        if (s < nchunks) issue_tma(s);
    }
    __syncthreads();
    for (int i = 0; i < nchunks; ++i) {
      const int s = i % STAGES;
      // refill the stage released one iteration ago (deferred so that thread 0 does not wait for the slowest warp)
      if (tid == 0 && i >= 1 && i - 1 + STAGES < nchunks) {
        mbar_wait(&empty_bar[(i - 1) % STAGES], ((i - 1) / STAGES) & 1);
        issue_tma(i - 1 + STAGES);
      }
      mbar_wait(&full_bar[s], (i / STAGES) & 1);
      compute(smem + s * STAGE);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
  } else {
    const float* f1b = f1 + (size_t)b * C * H * W;
    const float* f2b = f2 + (size_t)b * C * H * W;
    auto issue = [&](int i) {
      float* st = smem + (i % STAGES) * STAGE;
      const int c0 = (cr.begin + i) * CC;
      stage_box_async<CC, TH, TW, S1, T::THREADS, STG == STG_ASYNC16 ? 4 : 1>(st, f1b, c0, C, H, W, y0, x0);
      // the halo starts at x0 - D: 16-byte copies need D % 4 == 0 (d = 10 stages its halo with 8-byte copies)
      stage_box_async<CC, F2H, F2W, S2, T::THREADS, STG != STG_ASYNC16 ? 1 : (D % 4 == 0 ? 4 : 2)>(st + T::F1_STAGE, f2b, c0, C, H, W, y0 - D + dy0, x0 - D);
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < nchunks) issue(s);
      cp_async_commit();
    }
    for (int i = 0; i < nchunks; ++i) {
      cp_async_wait<STAGES - 2>();  // chunk i has landed (for this thread's copies) ...
      __syncthreads();              // ... and for everybody's; everybody is also done computing chunk i-1
      if (i + STAGES - 1 < nchunks) issue(i + STAGES - 1);
      cp_async_commit();
      compute(smem + (i % STAGES) * STAGE);
    }
    cp_async_wait<0>();
  }

  const size_t bstride = out_bstride ? (size_t)out_bstride : (size_t)ND * ND * H * W;
  constexpr bool VEC = STG != STG_ASYNC4;
  if (ksplit == 1) {
    const int y = y0 + ty, xs = x0 + tx * PX;
    if (y >= H || xs >= W) return;
    float* ob = out + (size_t)b * bstride + ((size_t)((dy0 + dyi) * ND) * H + y) * W + xs;
    const int Wb = (W + 7) >> 3;
#pragma unroll
    for (int dx = 0; dx < ND; ++dx) {
      float r[PX];
      unsigned mbyte = 0u;
#pragma unroll
      for (int p = 0; p < PX; ++p) {
        const float v = acc_get(dx, p) * inv_c;
        mbyte |= (v > 0.f ? 1u : 0u) << p;
        r[p] = v > 0.f ? v : v * slope;
      }
      if (mask != nullptr) mask[(((size_t)b * ND * ND + (dy0 + dyi) * ND + dx) * H + y) * Wb + (xs >> 3)] = (unsigned char)mbyte;
      float* o = ob + (size_t)dx * H * W;
      if (VEC) {  // W % 4 == 0 and 16B-aligned rows: each float4 is entirely inside or outside
#pragma unroll
        for (int q = 0; q < PX / 4; ++q)
          if (xs + 4 * q < W) *reinterpret_cast<float4*>(o + 4 * q) = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
      } else {
#pragma unroll
        for (int p = 0; p < PX; ++p)
          if (xs + p < W) o[p] = r[p];
      }
    }
    return;
  }

  // ---- channel-split: reduce the ksplit partial cost volumes of the cluster through DSMEM ----
  if constexpr (T::NDY != T::ND) {
    return;   // the dy-group tiles (d = 10) are never launched with a channel split
  } else {
  cg::cluster_group cluster = cg::this_cluster();
  __syncthreads();  // every warp is done reading the staging ring; reuse it as the partial tile [ND*ND][TH][S1]
  float* part = smem;
  static_assert(T::STAGES * STAGE >= ND * ND * TH * S1, "staging ring too small to hold the partial cost volume tile");
  {
#pragma unroll
    for (int dx = 0; dx < ND; ++dx) {
      float* pp = part + ((dyi * ND + dx) * TH + ty) * S1 + tx * PX;
#pragma unroll
      for (int q = 0; q < PX / 4; ++q)
        *reinterpret_cast<float4*>(pp + 4 * q) = make_float4(acc_get(dx, 4 * q), acc_get(dx, 4 * q + 1), acc_get(dx, 4 * q + 2), acc_get(dx, 4 * q + 3));
    }
  }
  cluster.sync();
  // rank r finalises planes k = r, r + ksplit, ...: all (plane, row, float4) items are spread over the whole CTA and the
  // ksplit peer reads of an item are issued together (independent DSMEM loads, ~215 cycles each)
  const unsigned rank = cluster.block_rank();
  constexpr int ROW4 = TW / 4, PLANE4 = TH * ROW4;
  static_assert(ROW4 % 2 == 0 && T::THREADS % 2 == 0, "mask nibbles are paired across adjacent lanes");
  const int nplanes = (ND * ND - (int)rank + ksplit - 1) / ksplit;
  const int Wb = (W + 7) >> 3;
  const float* peers[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) peers[q] = cluster.map_shared_rank(part, q < ksplit ? q : 0);
  const int nitems = nplanes * PLANE4;
  for (int it0 = 0; it0 < nitems; it0 += blockDim.x) {  // uniform trip count: every lane takes part in the shuffle below
    const bool valid = it0 + tid < nitems;
    const int it = valid ? it0 + tid : 0;
    const int kp = it / PLANE4, i = it - kp * PLANE4;
    const int k = (int)rank + kp * ksplit;
    const int ry = i / ROW4, x4 = i - ry * ROW4;
    const int off = (k * TH + ry) * S1 + x4 * 4;
    float4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < ksplit) v[q] = *reinterpret_cast<const float4*>(peers[q] + off);
    float4 s = v[0];
#pragma unroll
    for (int q = 1; q < 8; ++q)
      if (q < ksplit) { s.x += v[q].x; s.y += v[q].y; s.z += v[q].z; s.w += v[q].w; }
    const int y = y0 + ry, x = x0 + x4 * 4;
    float r[4] = {s.x * inv_c, s.y * inv_c, s.z * inv_c, s.w * inv_c};
    unsigned nib = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      nib |= (r[j] > 0.f ? 1u : 0u) << j;
      r[j] = r[j] > 0.f ? r[j] : r[j] * slope;
    }
    // items it (even) and it + 1 are the two halves of one 8-pixel mask byte and sit in adjacent lanes of one warp
    // (item count and loop stride are even), so the odd lane hands its nibble to the even one
    const unsigned other = __shfl_xor_sync(0xffffffffu, nib, 1);
    if (mask != nullptr && valid && (x4 & 1) == 0 && y < H && x < W)
      mask[(((size_t)b * ND * ND + k) * H + y) * Wb + (x >> 3)] = (unsigned char)(nib | (other << 4));
    if (valid && y < H && x < W) {
      float* o = out + (size_t)b * bstride + ((size_t)k * H + y) * W + x;
      if (VEC) {
        *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (x + j < W) o[j] = r[j];
      }
    }
  }
  cluster.sync();  // nobody may exit while a peer still reads its shared memory
  }
}

// ---- forward, persistent (regular shapes, no channel split) -----------------------------------------
// One CTA per resident slot (2 per SM) walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The TMA ring keeps
// running ACROSS tile boundaries: while the warps finish tile t and store its 81 planes, the boxes of tile t+1 are
// already in flight, so only the first tile of a CTA pays the cold-start latency (at C = 32 a tile is just 4 chunks --
// without this every tile would wait ~1.5 us for its first box with nothing else to do).
// TMAST: the epilogue goes through shared memory and TMA stores.  Every dy-warp owns a [9 planes][TH][TW] staging box; after a
// tile's main loop it waits until the TMA unit has read its previous box, writes the activated values (128-bit shared stores),
// and lane 0 issues ONE cp.async.bulk.tensor store for the 9 planes -- no per-thread global stores, no bounds predicates (the TMA
// unit clips at the image border), and the copy-out runs while the warp is already in the next tile's main loop.
template <class T, int UNROLL, bool TMAST = false>
__global__ void __maxnreg__(OCF_FWD_REGS)
corr_fwd_persist(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap mapo, float* __restrict__ out,
                 unsigned char* __restrict__ mask, int C, int H, int W, long long out_bstride, float inv_c, float slope, int tiles_x,
                 int tiles_y, int ntiles) {
  constexpr int D = T::D, PX = T::PX, ND = T::ND, TH = T::TH, TW = T::TW, CC = T::CC, STAGES = T::STAGES;
  constexpr int S1 = T::S1, S2 = T::S2, F2H = T::F2H, WIN = T::WIN;
  constexpr int STAGE = T::F1_STAGE + T::F2_STAGE;
  constexpr unsigned BYTES = sizeof(float) * STAGE;
  static_assert(PX % 2 == 0, "pixel pairs");
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) unsigned long long full_bar[STAGES], empty_bar[STAGES];

  const int tid = threadIdx.x;
  const int lane = tid % T::LANES;
  const int tx = lane % T::TXT, ty = lane / T::TXT;
  const int dyi = tid / T::LANES;
  OCF_TL(0);
  OCF_TL_SM();
  const int nchunks = (C + CC - 1) / CC;
  const int ntl = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int total = ntl * nchunks;
  const size_t bstride = out_bstride ? (size_t)out_bstride : (size_t)ND * ND * H * W;

  // producer state (thread 0 only): next flat chunk to issue = (tile iteration pk, chunk pi)
  int pk = 0, pi = 0, pn = 0;
  auto issue_next = [&]() {
    const int tile = (int)blockIdx.x + pk * (int)gridDim.x;
    const int txi = tile % tiles_x, r = tile / tiles_x;
    const int tyi = r % tiles_y, b = r / tiles_y;
    const int s = pn % STAGES;
    float* st = smem + s * STAGE;
    mbar_expect_tx(&full_bar[s], BYTES);
    tma_load_4d(st, &map1, &full_bar[s], txi * TW, tyi * TH, pi * CC, b);
    tma_load_4d(st + T::F1_STAGE, &map2, &full_bar[s], txi * TW - D, tyi * TH - D, pi * CC, b);
    ++pn;
    if (++pi == nchunks) { pi = 0; ++pk; }
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], ND); }
    mbar_fence_init();
  }
  __syncthreads();
  if (dyi == ND) {
    // dedicated producer warp (the CTA is alone on its SM, so the extra warp is free): the compute warps never wait for a stage
    // to drain -- when thread 0 doubled as the producer, warp 0 stalled on the slowest warp once per chunk
    if (lane == 0) {
      for (int n = 0; n < total; ++n) {
        if (n >= STAGES) mbar_wait(&empty_bar[n % STAGES], ((n / STAGES) - 1) & 1);
        issue_next();
      }
    }
    return;
  }

  u64 accp[PX][D];
  float accs[PX];
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    accs[p] = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) accp[p][j] = 0ull;
  }
  const int o1 = ty * S1 + tx * PX;
  const int o2 = T::F1_STAGE + (ty + dyi) * S2 + tx * PX;

  int n = 0;  // flat chunk counter of this CTA
  for (int k = 0; k < ntl; ++k) {
    for (int i = 0; i < nchunks; ++i, ++n) {
      const int s = n % STAGES;
      mbar_wait(&full_bar[s], (n / STAGES) & 1);
      if (n == 0) OCF_TL(1);
      const float* st = smem + s * STAGE;
      const float* p1 = st + o1;
      const float* p2 = st + o2;
#pragma unroll UNROLL
      for (int c = 0; c < CC; ++c) {
        float a[PX];
        u64 w2[WIN / 2];
#pragma unroll
        for (int q = 0; q < PX / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(p1 + c * TH * S1 + 4 * q);
          a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < WIN / 4; ++q) {
          const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p2 + c * F2H * S2 + 4 * q);
          w2[2 * q] = v.x; w2[2 * q + 1] = v.y;
        }
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          const u64 ap = pack2(a[p], a[p]);
          const int first = p & 1;
#pragma unroll
          for (int j = 0; j < D; ++j) ffma2(accp[p][j], ap, w2[(p + first + 2 * j) / 2]);
          float lo, hi;
          if (first == 0) {
            unpack2(w2[(p + 2 * D) / 2], lo, hi);
            accs[p] = fmaf(a[p], lo, accs[p]);
          } else {
            unpack2(w2[p / 2], lo, hi);
            accs[p] = fmaf(a[p], hi, accs[p]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    // ---- epilogue of tile k: 1/C, LeakyReLU, 128-bit stores; accumulators restart from zero ----
    if (k < 3) OCF_TL(2 + 2 * k);
    const int tile = (int)blockIdx.x + k * (int)gridDim.x;
    const int txi = tile % tiles_x, r = tile / tiles_x;
    const int tyi = r % tiles_y, b = r / tiles_y;
    const int y = tyi * TH + ty, xs = txi * TW + tx * PX;
    const bool inside = y < H && xs < W;
    float* ob = out + (size_t)b * bstride + ((size_t)(dyi * ND) * H + y) * W + xs;
    const int Wb = (W + 7) >> 3;
    float* stg = smem + STAGES * STAGE + dyi * (ND * TH * TW);   // TMAST: this warp's [ND][TH][TW] staging box
    if constexpr (TMAST) {
      if (lane == 0) tma_store_wait_read();   // the previous tile's box of this warp has been read by the TMA unit
      __syncwarp();
    }
#pragma unroll
    for (int dx = 0; dx < ND; ++dx) {
      float rr[PX];
      unsigned mbyte = 0u;
#pragma unroll
      for (int p = 0; p < PX; ++p) {
        const int first = p & 1;
        float v;
        if ((first == 0 && dx == 2 * D) || (first == 1 && dx == 0)) {
          v = accs[p];
        } else {
          float lo, hi;
          unpack2(accp[p][(dx - first) / 2], lo, hi);
          v = ((dx - first) & 1) ? hi : lo;
        }
        v *= inv_c;
        mbyte |= (v > 0.f ? 1u : 0u) << p;
        rr[p] = v > 0.f ? v : v * slope;
      }
      if (inside && mask != nullptr) mask[(((size_t)b * ND * ND + dyi * ND + dx) * H + y) * Wb + (xs >> 3)] = (unsigned char)mbyte;
      if constexpr (TMAST) {
        float* sp = stg + (dx * TH + ty) * TW + tx * PX;
#pragma unroll
        for (int q = 0; q < PX / 4; ++q) *reinterpret_cast<float4*>(sp + 4 * q) = make_float4(rr[4 * q], rr[4 * q + 1], rr[4 * q + 2], rr[4 * q + 3]);
      } else if (inside) {
        float* o = ob + (size_t)dx * H * W;
#pragma unroll
        for (int q = 0; q < PX / 4; ++q)
          if (xs + 4 * q < W) *reinterpret_cast<float4*>(o + 4 * q) = make_float4(rr[4 * q], rr[4 * q + 1], rr[4 * q + 2], rr[4 * q + 3]);
      }
    }
    if constexpr (TMAST) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the TMA unit
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(stg, &mapo, txi * TW, tyi * TH, dyi * ND, b);
        tma_store_commit();
      }
    }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      accs[p] = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) accp[p][j] = 0ull;
    }
    if (k < 3) OCF_TL(3 + 2 * k);
  }
  if constexpr (TMAST) {
    if (lane == 0) tma_store_wait_all();   // the stores must have completed before the CTA (and its shared memory) goes away
  }
  OCF_TL(14);
}

// Any displacement up to OCF_MAX_DISPLACEMENT: one thread per (pixel, dy), operands through L1/L2.
__global__ void __launch_bounds__(128)
corr_fwd_generic(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ out, int C, int H, int W,
                 int d, long long out_bstride, float inv_c, float slope, const float* __restrict__ norm) {
  constexpr int NDMAX = 2 * OCF_MAX_DISPLACEMENT + 1;
  const int nd = 2 * d + 1;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= H * W) return;
  const int y = pix / W, x = pix - y * W;
  const int dyi = blockIdx.y, b = blockIdx.z;
  const int yy = y + dyi - d;
  float nmean = 0.f, ninv = 1.f;
  if (norm) { nmean = __ldg(norm); ninv = __ldg(norm + 1); }
  float acc[NDMAX];
#pragma unroll
  for (int i = 0; i < NDMAX; ++i) acc[i] = 0.f;
  if (yy >= 0 && yy < H) {
    const float* p1 = f1 + ((size_t)b * C * H + y) * W + x;
    const float* p2 = f2 + ((size_t)b * C * H + yy) * W;
    for (int c = 0; c < C; ++c) {
      const float a = (__ldg(p1 + (size_t)c * H * W) - nmean) * ninv;
#pragma unroll
      for (int i = 0; i < NDMAX; ++i) {
        const int xx = x + i - d;
        if (i < nd && xx >= 0 && xx < W) acc[i] = fmaf(a, (__ldg(p2 + (size_t)c * H * W + xx) - nmean) * ninv, acc[i]);
      }
    }
  }
  const size_t bstride = out_bstride ? (size_t)out_bstride : (size_t)nd * nd * H * W;
  float* o = out + (size_t)b * bstride + ((size_t)(dyi * nd) * H + y) * W + x;
#pragma unroll
  for (int i = 0; i < NDMAX; ++i)
    if (i < nd) {
      const float v = acc[i] * inv_c;
      o[(size_t)i * H * W] = v > 0.f ? v : v * slope;
    }
}

// ---- backward -----------------------------------------------------------------------------------
// mode 0: dout = d f1, fo = f2 ; mode 1: dout = d f2, fo = f1.
// blockIdx.z = (b * nmodes + slot) * ksplit + channel slice.
// GD (TMA staging only, experiments): 1 = the 81 coefficient planes go global -> registers with 128-bit loads instead of through
// a TMA-staged shared-memory copy, so that the feature ring starts filling at kernel entry; 2 = the same for mode 0 only (d f1:
// aligned planes, 18 loads straight into their final registers), mode 1 keeps the staged lift (unaligned: 27 loads through 12
// temporaries per displacement).  pf_stride > 0: the producer thread asks the TMA unit to PREFETCH INTO L2 the coefficient
// boxes and the first feature boxes of the CTA pf_stride linear block ids ahead (the one that will take over a slot of this
// wave), once its own ring is full -- the next wave's prologue then reads L2 instead of DRAM.  3 = EARLY first stage (use with a
// CC = 4 tile): the coefficient boxes are 40 instead of 44 floats wide (2-way bank conflicts in the one-off lift), which leaves
// room for ONE feature stage next to the staging area, so the first feature box travels together with the coefficients instead
// of one more memory round trip behind them; the other stages and the reduction buffer alias the staging area as before.
// Measured (see the launcher): only the prefetch pays; it is on by default, GD stays 0.
template <class T, int CR, int STG, int GD = 0>
__global__ void __launch_bounds__(Threads<T, STG>::value, 2)
corr_bwd_tiled(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
               const __grid_constant__ CUtensorMap mapg, const float* __restrict__ g,
               const float* __restrict__ oact, const unsigned char* __restrict__ mask, const float* __restrict__ f1,
               const float* __restrict__ f2, float* __restrict__ df1,
               float* __restrict__ df2, int C, int H, int W, long long g_bstride, long long a_bstride, float inv_c, float slope,
               int nmodes, int first_mode, int ksplit, long long f1_bstride, long long f2_bstride, int pf_stride) {
  constexpr int D = T::D, PX = T::PX, ND = T::ND, TH = T::TH, TW = T::TW, CC = T::CC, STAGES = T::STAGES;
  constexpr int S1 = T::S1, S2 = T::S2, F2H = T::F2H, F2W = T::F2W, WIN = T::WIN;
  constexpr bool TMA = STG == STG_TMA;
  constexpr bool VEC = STG != STG_ASYNC4;
  static_assert(CC % CR == 0, "CC must be a multiple of CR");
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) unsigned long long full_bar[STAGES], empty_bar[STAGES], g_bar[T::ND], gdone_bar;   // g_bar: one per dy box
  constexpr bool EARLY = GD == 3;
  static_assert(!EARLY || (STG == STG_TMA && T::NGY == 1), "the early first stage is a TMA-path layout");
  float* red = smem + (EARLY ? STAGES - 1 : STAGES) * T::F2_STAGE;  // [NDY][CR][TH][S1]
  // TMA variant: the 81 coefficient planes of the tile are staged through shared memory first (one {GW, TH, ND} box per
  // dy-warp, GW == 12 mod 32 floats so the 128-bit reads are conflict-free); the area is then reused by the ring + red.
  constexpr int GW = EARLY ? T::F2W : T::S2, BOXG = ND * TH * GW;
  // ring stage s: EARLY keeps stage 0 behind the staging area (never aliased), stages 1.. at the bottom
  auto stage_base = [&](int s) -> float* {
    if constexpr (EARLY) return s == 0 ? smem + ND * BOXG : smem + (s - 1) * T::F2_STAGE;
    else return smem + s * T::F2_STAGE;
  };

  const int tid = threadIdx.x;
  const int lane = tid % T::LANES;
  const int tx = lane % T::TXT, ty = lane / T::TXT;
  const int dyi = tid / T::LANES;  // == ND for the TMA producer warp
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  // blockIdx.z = ((b * nmodes + slot) * NGY + dy group) * ksplit + channel slice.  With several dy groups per tile (d = 10) every
  // group adds its partial sum into the zero-filled output with vector reds; one group (d = 4) stores.
  const int zz0 = blockIdx.z / ksplit, ks = blockIdx.z - zz0 * ksplit;
  const int zz = zz0 / T::NGY, dy0 = (zz0 - zz * T::NGY) * T::NDY;
  static_assert(T::NGY == 1 || STG != STG_TMA, "the TMA-staged backward handles all displacements in one CTA");
  const int b = zz / nmodes;
  const int mode = first_mode + (zz - b * nmodes);
  const bool direct = GD == 1 || (GD == 2 && mode == 0);   // compile-time constant unless GD == 2
  const ChunkRange cr = chunk_range(C, CC, ksplit, ks);
  const int nchunks = cr.count;
  float* dout = (mode == 0 ? df1 : df2) + (size_t)b * C * H * W;
  const size_t gb = (g_bstride ? (size_t)g_bstride : (size_t)ND * ND * H * W) * b;
  const size_t ab = (a_bstride ? (size_t)a_bstride : (size_t)ND * ND * H * W) * b;

  float G[ND][PX];  // 81 per-pixel coefficients of this thread's dy row, kept in registers for the whole kernel
  if constexpr (TMA) {
    OCF_TLB(tid == 0, 0);
    OCF_TLB_VAL(tid == 0, 5, mode + 1);
#ifdef OCF_TIMELINE
    if (tid == 0) { unsigned sm_; asm("mov.u32 %0, %smid;" : "=r"(sm_)); OCF_TLB_VAL(true, 15, sm_); }
#endif
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], ND); }
#pragma unroll
      for (int w = 0; w < ND; ++w) mbar_init(&g_bar[w], 1);
      mbar_init(&gdone_bar, ND);
      mbar_fence_init();
    }
    __syncthreads();
    const bool has_act = oact != nullptr || mask != nullptr;
    if (dyi == ND) {
      if (lane == 0) {
        if (!direct) {
        // coefficient boxes: mode 0 -> planes dy*ND.. at (y0, x0) ; mode 1 -> planes (2D-dy)*ND.. at (y0+dy-D, x0-D)
        // (one barrier per box: a warp lifts its nine planes as soon as ITS box has landed, while the later boxes are still
        // on their way, instead of all nine warps waiting for the last byte of the 114 KB)
        constexpr unsigned GBYTES = sizeof(float) * BOXG;
        for (int w = 0; w < ND; ++w) {
          const int plane0 = mode == 0 ? w * ND : (2 * D - w) * ND;
          unsigned long long* gb_ = &g_bar[OCF_BWD_GBAR_PER_BOX ? w : 0];
          if (OCF_BWD_GBAR_PER_BOX || w == 0) mbar_expect_tx(gb_, OCF_BWD_GBAR_PER_BOX ? GBYTES : GBYTES * ND);
          tma_load_4d(smem + w * BOXG, &mapg, gb_, mode == 0 ? x0 : x0 - D, mode == 0 ? y0 : y0 + w - D, plane0, b);
        }
        if constexpr (EARLY) {
          if (nchunks > 0) {   // the first feature box rides along with the coefficients (its stage is not part of the staging area)
            mbar_expect_tx(&full_bar[0], (unsigned)sizeof(float) * T::F2_STAGE);
            tma_load_4d(stage_base(0), mode == 0 ? &map2 : &map1, &full_bar[0], x0 - D, y0 - D, cr.begin * CC, b);
          }
        }
        mbar_wait(&gdone_bar, 0);  // every warp has lifted its coefficients out of the staging area: feed the feature ring
        }
        constexpr unsigned BYTES = sizeof(float) * T::F2_STAGE;
        const CUtensorMap* map = mode == 0 ? &map2 : &map1;  // the OTHER feature
        const int pf_at = min(nchunks, STAGES) - 1;          // the ring is full after this chunk: time to think of the next wave
        for (int i = EARLY ? 1 : 0; i < nchunks; ++i) {
          const int s = i % STAGES;
          if (i >= STAGES) mbar_wait(&empty_bar[s], ((i / STAGES) - 1) & 1);
          mbar_expect_tx(&full_bar[s], BYTES);
          tma_load_4d(stage_base(s), map, &full_bar[s], x0 - D, y0 - D, (cr.begin + i) * CC, b);
          if (pf_stride > 0 && i == pf_at) {
            // L2 prefetch for the CTA pf_stride linear ids ahead: same decode as above, hints only (nothing waits on them)
            const long long lin = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z) + pf_stride;
            const long long per_z = (long long)gridDim.x * gridDim.y;
            if (lin < per_z * gridDim.z) {
              const int nz = (int)(lin / per_z), rem = (int)(lin - nz * per_z);
              const int nyb = rem / (int)gridDim.x, nxb = rem - nyb * (int)gridDim.x;
              const int nzz0 = nz / ksplit, nks = nz - nzz0 * ksplit;
              const int nzz = nzz0 / T::NGY;
              const int nb = nzz / nmodes, nmode = first_mode + (nzz - nb * nmodes);
              const int nx0 = nxb * TW, ny0 = nyb * TH;
              for (int w = 0; w < ND; ++w)
                tma_prefetch_4d(&mapg, nmode == 0 ? nx0 : nx0 - D, nmode == 0 ? ny0 : ny0 + w - D, nmode == 0 ? w * ND : (2 * D - w) * ND, nb);
              const ChunkRange ncr = chunk_range(C, CC, ksplit, nks);
              const CUtensorMap* nmap = nmode == 0 ? &map2 : &map1;
              for (int j = 0; j < min(ncr.count, STAGES); ++j) tma_prefetch_4d(nmap, nx0 - D, ny0 - D, (ncr.begin + j) * CC, nb);
            }
          }
        }
      }
      return;  // the producer warp takes no part in the compute-warp barriers below
    }
    // LeakyReLU mask: the sign of the activated forward output at this thread's 81 coefficient positions, fetched straight
    // from global memory into a 72-bit register mask WHILE the coefficient boxes are in flight (a second TMA pass over the
    // staging area would serialise another L2 round trip + lift behind the first).  Where the coefficient itself is a
    // zero-filled out-of-image element the mask value is irrelevant, so addresses are clamped instead of predicated
    // (W % 4 == 0 on this path: an aligned group of 4 is entirely inside or entirely outside the row).
    unsigned mbits[3] = {0u, 0u, 0u};  // bit (dx % 3) * PX + p of word dx / 3
    static_assert(ND == 9 && PX == 8, "mask packing assumes 9 displacements x 8 pixels");
    if (mask != nullptr) {
      // the forward's sign bitmask (1 bit per cost-volume element, 8 pixels per byte): 9 / 18 byte loads per thread
      const int Wb = (W + 7) >> 3;
      const unsigned char* mp = mask + (size_t)b * ND * ND * H * Wb;
      const int xs = x0 + tx * PX;
      if (mode == 0) {
        const int y = min(y0 + ty, H - 1), bx = min(xs >> 3, Wb - 1);
#pragma unroll
        for (int dx = 0; dx < ND; ++dx)
          mbits[dx / 3] |= (unsigned)__ldg(mp + ((size_t)(dyi * ND + dx) * H + y) * Wb + bx) << ((dx % 3) * PX);
      } else {
        const int sy = min(max(y0 + ty + dyi - D, 0), H - 1);
#pragma unroll
        for (int dx = 0; dx < ND; ++dx) {
          const unsigned char* row = mp + ((size_t)((2 * D - dyi) * ND + (2 * D - dx)) * H + sy) * Wb;
          const int c0 = xs + dx - D;        // first needed column (>= -D); out-of-image columns carry zero coefficients
          const int fb = c0 >> 3;            // floor(c0 / 8), -1 at the left border
          const unsigned lo = __ldg(row + min(max(fb, 0), Wb - 1)), hi = __ldg(row + min(fb + 1, Wb - 1));
          mbits[dx / 3] |= (((lo | (hi << 8)) >> (c0 & 7)) & 0xffu) << ((dx % 3) * PX);
        }
      }
    } else if (oact != nullptr) {
      const float* ap = oact + ab;
      const int xs = x0 + tx * PX;
      if (mode == 0) {
        const int y = min(y0 + ty, H - 1);
#pragma unroll
        for (int dx = 0; dx < ND; ++dx) {
          const float* row = ap + ((size_t)(dyi * ND + dx) * H + y) * W;
#pragma unroll
          for (int q = 0; q < PX / 4; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(row + min(xs + 4 * q, W - 4)));
            const unsigned m = (v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) | (v.w > 0.f ? 8u : 0u);
            mbits[dx / 3] |= m << ((dx % 3) * PX + 4 * q);
          }
        }
      } else {
        const int sy = min(max(y0 + ty + dyi - D, 0), H - 1);
#pragma unroll
        for (int dx = 0; dx < ND; ++dx) {
          const float* row = ap + ((size_t)((2 * D - dyi) * ND + (2 * D - dx)) * H + sy) * W;
          const int cb = xs - D + (dx & ~3);  // aligned group holding the first needed column xs + dx - D
          unsigned m = 0u;                    // 12 sign bits of columns cb .. cb + 11
#pragma unroll
          for (int q = 0; q < PX / 4 + 1; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(row + min(max(cb + 4 * q, 0), W - 4)));
            m |= ((v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) | (v.w > 0.f ? 8u : 0u)) << (4 * q);
          }
          mbits[dx / 3] |= ((m >> (dx & 3)) & 0xffu) << ((dx % 3) * PX);
        }
      }
    }
    if (direct) {
      // W % 4 == 0 on this path: an aligned group of 4 columns is entirely inside or entirely outside the row
      const float* gp = g + gb;
      const int xs = x0 + tx * PX;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mode == 0) {
        const int y = y0 + ty;
#pragma unroll
        for (int dx = 0; dx < ND; ++dx) {
          const float* row = gp + ((size_t)(dyi * ND + dx) * H + min(y, H - 1)) * W;
#pragma unroll
          for (int q = 0; q < PX / 4; ++q) {
            const int x = xs + 4 * q;
            const float4 v = (y < H && x < W) ? __ldg(reinterpret_cast<const float4*>(row + x)) : z4;
            G[dx][4 * q] = v.x; G[dx][4 * q + 1] = v.y; G[dx][4 * q + 2] = v.z; G[dx][4 * q + 3] = v.w;
          }
        }
      } else {
        const int sy = y0 + ty + dyi - D;
        const bool yok = sy >= 0 && sy < H;
#pragma unroll
        for (int dx = 0; dx < ND; ++dx) {
          const float* row = gp + ((size_t)((2 * D - dyi) * ND + (2 * D - dx)) * H + min(max(sy, 0), H - 1)) * W;
          const int cb = xs - D + (dx & ~3);   // aligned group holding the first needed column xs + dx - D
          float t[PX + 4];
#pragma unroll
          for (int q = 0; q < PX / 4 + 1; ++q) {
            const int x = cb + 4 * q;
            const bool need = q < PX / 4 || (dx & 3) != 0;   // the third group only when the window is not aligned
            const float4 v = (need && yok && x >= 0 && x < W) ? __ldg(reinterpret_cast<const float4*>(row + x)) : z4;
            t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
          }
#pragma unroll
          for (int p = 0; p < PX; ++p) G[dx][p] = t[p + (dx & 3)];
        }
      }
    } else {
    const float* gw = smem + dyi * BOXG + ty * GW + tx * PX;
    mbar_wait(&g_bar[OCF_BWD_GBAR_PER_BOX ? dyi : 0], 0);
    OCF_TLB(tid == 0, 1);
    if (mode == 0) {
#pragma unroll
      for (int dx = 0; dx < ND; ++dx) {
#pragma unroll
        for (int q = 0; q < PX / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(gw + dx * TH * GW + 4 * q);
          G[dx][4 * q] = v.x; G[dx][4 * q + 1] = v.y; G[dx][4 * q + 2] = v.z; G[dx][4 * q + 3] = v.w;
        }
      }
    } else {
#pragma unroll
      for (int dx = 0; dx < ND; ++dx) {
        float t[PX + 4];
#pragma unroll
        for (int q = 0; q < PX / 4 + 1; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(gw + (2 * D - dx) * TH * GW + (dx & ~3) + 4 * q);
          t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int p = 0; p < PX; ++p) G[dx][p] = t[p + (dx & 3)];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&gdone_bar);
    // EARLY: the first feature stage is already there, so nothing else keeps a fast warp from writing its partial sums into
    // the reduction buffer -- which aliases the boxes slower warps are still lifting
    if constexpr (EARLY) mbar_wait(&gdone_bar, 0);
    }
    if (has_act) {
#pragma unroll
      for (int dx = 0; dx < ND; ++dx)
#pragma unroll
        for (int p = 0; p < PX; ++p)
          if (!((mbits[dx / 3] >> ((dx % 3) * PX + p)) & 1u)) G[dx][p] *= slope;
    }
    OCF_TLB(tid == 0, 2);
  }
  const float* fo = (mode == 0 ? f2 + (size_t)b * (f2_bstride ? (size_t)f2_bstride : (size_t)C * H * W)
                               : f1 + (size_t)b * (f1_bstride ? (size_t)f1_bstride : (size_t)C * H * W));
  auto issue = [&](int i) {
    stage_box_async<CC, F2H, F2W, S2, T::THREADS, STG != STG_ASYNC16 ? 1 : (D % 4 == 0 ? 4 : 2)>(smem + (i % STAGES) * T::F2_STAGE, fo, (cr.begin + i) * CC,
                                                                                                  C, H, W, y0 - D + dy0, x0 - D);
  };
  if constexpr (!TMA) {
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < nchunks) issue(s);
      cp_async_commit();
    }
    const int y = y0 + ty, xs = x0 + tx * PX;
#pragma unroll
    for (int dx = 0; dx < ND; ++dx) {
      // mode 0: plane k(dy,dx) at (y, x) ; mode 1: plane k(-dy,-dx) at (y+dy, x+dx)
      const int dyg = dy0 + dyi;   // this warp's vertical displacement index in 0 .. ND-1
      const int k = mode == 0 ? dyg * ND + dx : (2 * D - dyg) * ND + (2 * D - dx);
      const int sy = mode == 0 ? y : y + dyg - D;
      const int sx0 = mode == 0 ? xs : xs + dx - D;
      const size_t off = gb + ((size_t)k * H + sy) * W;
#pragma unroll
      for (int p = 0; p < PX; ++p) {
        const int sx = sx0 + p;
        float v = 0.f;
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
          v = __ldg(g + off + sx);
          if (mask != nullptr) {
            const int Wb = (W + 7) >> 3;
            if (!((__ldg(mask + (((size_t)b * ND * ND + k) * H + sy) * Wb + (sx >> 3)) >> (sx & 7)) & 1)) v *= slope;
          } else if (oact != nullptr && !(__ldg(oact + (off - gb + ab) + sx) > 0.f)) v *= slope;
        }
        G[dx][p] = v;
      }
    }
  }

  for (int i = 0; i < nchunks; ++i) {
    const int s = i % STAGES;
    if constexpr (TMA) {
      mbar_wait(&full_bar[s], (i / STAGES) & 1);
      OCF_TLB(tid == 0 && i == 0, 3);
    } else {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      if (i + STAGES - 1 < nchunks) issue(i + STAGES - 1);
      cp_async_commit();
    }
    const int c0 = (cr.begin + i) * CC;
    const float* pw = stage_base(s) + (ty + dyi) * S2 + tx * PX;
#pragma unroll 1
    for (int r0 = 0; r0 < CC; r0 += CR) {
      if (c0 + r0 >= C) break;  // uniform across the block
#pragma unroll
      for (int c = 0; c < CR; ++c) {
        float w[WIN], part[PX];
#pragma unroll
        for (int q = 0; q < WIN / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(pw + (r0 + c) * F2H * S2 + 4 * q);
          w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int p = 0; p < PX; ++p) part[p] = 0.f;
        // (Measured: the same loop as packed fma.rn.f32x2 over pixel pairs -- aligned window pairs for even dx, a shifted copy of
        // the window for odd dx -- is 8-15 % SLOWER (L2 level 65.6 vs 60.4 us, C = 128: 208 vs 181 us): the 15 register-pair
        // packs per channel cost more issue slots than the 36 saved FMAs.)
#pragma unroll
        for (int dx = 0; dx < ND; ++dx)
#pragma unroll
          for (int p = 0; p < PX; ++p) part[p] = fmaf(G[dx][p], w[p + dx], part[p]);
        float* rp = red + ((dyi * CR + c) * TH + ty) * S1 + tx * PX;
#pragma unroll
        for (int q = 0; q < PX / 4; ++q)
          *reinterpret_cast<float4*>(rp + 4 * q) = make_float4(part[4 * q], part[4 * q + 1], part[4 * q + 2], part[4 * q + 3]);
      }
      if constexpr (TMA) consumer_bar_sync<T::THREADS>(); else __syncthreads();
      // cross-dy reduction: CR*TH*TW/4 float4 outputs
      constexpr int OUT4 = CR * TH * TW / 4;
      if constexpr (TMA && T::NGY == 1 && OUT4 <= T::THREADS && OCF_BWD_FASTRED) {
        // TMA path (W % 4 == 0, dense 16-byte aligned outputs): one float4 per thread, 32-bit index math (the compiler
        // re-derives the indices in every group -- the 96 registers are all taken during the compute phase -- and did so in
        // 64 bits: ~50 of the ~105 instructions of this phase) and packed adds (add.f32x2: 16 instead of 32): 105 -> 73
        // instructions.  Same order of additions as the general form below (bit-identical results).
        if (tid < OUT4) {
          constexpr int X4 = TW / 4;
          const int x4 = tid % X4, ry = (tid / X4) % TH, c = tid / (X4 * TH);
          const float* rp = red + (c * TH + ry) * S1 + x4 * 4;
          float4 v = *reinterpret_cast<const float4*>(rp);
          u64 s01 = pack2(v.x, v.y), s23 = pack2(v.z, v.w);
#pragma unroll
          for (int k = 1; k < T::NDY; ++k) {
            v = *reinterpret_cast<const float4*>(rp + k * CR * TH * S1);
            fadd2(s01, pack2(v.x, v.y));
            fadd2(s23, pack2(v.z, v.w));
          }
          const int ch = c0 + r0 + c, y = y0 + ry, x = x0 + x4 * 4;
          if (ch < C && y < H && x < W) {
            const u64 ic = pack2(inv_c, inv_c);
            float4 o4;
            unpack2(fmul2(s01, ic), o4.x, o4.y);
            unpack2(fmul2(s23, ic), o4.z, o4.w);
            // C * H * W < 2^31 (checked by the launcher).  (Measured: parking the base pointer in shared memory -- the compiler
            // re-derives the 64-bit b * C * H * W in every group -- saves 14 instructions and is 4 % SLOWER: an LDS.64 in front of
            // a generic store on the critical path to the barrier.)
            *reinterpret_cast<float4*>(dout + (unsigned)((ch * H + y) * W + x)) = o4;
          }
        }
      } else
      for (int j = tid; j < OUT4; j += T::THREADS) {
        const int x4 = j % (TW / 4), row = j / (TW / 4);  // row = c*TH + ry
        const int c = row / TH, ry = row - c * TH;
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < T::NDY; ++k) {
          const float4 v = *reinterpret_cast<const float4*>(red + ((k * CR + c) * TH + ry) * S1 + x4 * 4);
          sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
        const int ch = c0 + r0 + c, y = y0 + ry, x = x0 + x4 * 4;
        if (ch < C && y < H && x < W) {
          float* o = dout + ((size_t)ch * H + y) * W + x;
          if (T::NGY > 1) {   // partial sum of one dy group: accumulate (the entry point zero-fills the output)
            if (VEC) red_add_v4(o, sum.x * inv_c, sum.y * inv_c, sum.z * inv_c, sum.w * inv_c);
            else {
              atomicAdd(o, sum.x * inv_c);
              if (x + 1 < W) atomicAdd(o + 1, sum.y * inv_c);
              if (x + 2 < W) atomicAdd(o + 2, sum.z * inv_c);
              if (x + 3 < W) atomicAdd(o + 3, sum.w * inv_c);
            }
          } else if (VEC) {
            *reinterpret_cast<float4*>(o) = make_float4(sum.x * inv_c, sum.y * inv_c, sum.z * inv_c, sum.w * inv_c);
          } else {
            o[0] = sum.x * inv_c;
            if (x + 1 < W) o[1] = sum.y * inv_c;
            if (x + 2 < W) o[2] = sum.z * inv_c;
            if (x + 3 < W) o[3] = sum.w * inv_c;
          }
        }
      }
      if constexpr (TMA) consumer_bar_sync<T::THREADS>(); else __syncthreads();
      OCF_TLB(TMA && tid == 0 && i == 0 && r0 == 0, 4);
    }
    if constexpr (TMA) {
      // the last consumer barrier above ordered every warp's reads of this stage: one arrival per warp frees it
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
  }
  if constexpr (!TMA) cp_async_wait<0>();
  OCF_TLB(TMA && tid == 0, 14);
}

// ---- backward, channel-split persistent form (d = 4, TMA, sign bitmask or no activation) ------------------------------
// The tiled backward above gives every vertical displacement its own warp and reduces the nine partial sums through
// shared memory with two block barriers per 4 channels.  Here a warp owns a GROUP OF CHANNELS instead and walks all nine
// dy itself, so nothing is reduced across warps and the main loop has no barrier at all:
//   * one persistent CTA per SM walks the work items (tile, mode);
//   * the 81 coefficient planes of the item are TMA-staged once into shared memory (114 KB, resident for the whole item), the
//     LeakyReLU sign mask is folded into the staged copy by a cooperative pass (one block barrier per item);
//   * compute warp w takes the 8-channel groups w, w + NW, ...: per dy it lifts its 72 coefficients (9 dx x 8 pixels) from
//     shared memory ONCE and applies them to the 8 channels of the group (8 x 72 FMAs against 18 + 32 LDS.128), accumulating
//     8 x 8 outputs in registers; the feature boxes of a group arrive through a TMA ring whose stages are per warp;
//   * outputs go straight from registers to global memory (128-bit stores).
template <class T, int NW, int RING>
__global__ void __launch_bounds__((NW + 1) * 32, 1)
corr_bwd_cs(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap mapg,
            const unsigned char* __restrict__ mask, float* __restrict__ df1, float* __restrict__ df2, int C, int H, int W, float inv_c,
            float slope, int nmodes, int first_mode, int tiles_x, int tiles_y, int nitems) {
  constexpr int D = T::D, PX = T::PX, ND = T::ND, TH = T::TH, TW = T::TW, CC = T::CC;
  constexpr int S2 = T::S2, F2H = T::F2H, WIN = T::WIN;
  constexpr int GW = T::S2, BOXG = ND * TH * GW;          // one {GW, TH, ND} box per dy
  constexpr int GFLOATS = ND * BOXG;
  static_assert(ND == 9 && PX == 8 && CC == 8, "d = 4 tile");
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) unsigned long long full_bar[RING], empty_bar[RING], g_bar, gfree_bar;
  float* gs = smem;                      // [dy][dx][TH][GW]
  float* ring = smem + GFLOATS;          // RING x F2_STAGE

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tx = lane % T::TXT, ty = lane / T::TXT;
  const int ngroups = (C + CC - 1) / CC;
  const int nit = ((int)blockIdx.x < nitems) ? (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < RING; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&g_bar, 1);
    mbar_init(&gfree_bar, NW);
    mbar_fence_init();
  }
  __syncthreads();

  auto item_coord = [&](int it, int& b, int& mode, int& x0, int& y0) {
    const int item = (int)blockIdx.x + it * (int)gridDim.x;
    const int tile = item / nmodes;
    mode = first_mode + (item - tile * nmodes);
    const int txi = tile % tiles_x, r = tile / tiles_x;
    x0 = txi * TW; y0 = (r % tiles_y) * TH; b = r / tiles_y;
  };

  if (warp == NW) {
    // =========================== producer ===========================
    if (lane == 0) {
      int n = 0;   // flat feature-stage counter
      for (int it = 0; it < nit; ++it) {
        int b, mode, x0, y0;
        item_coord(it, b, mode, x0, y0);
        if (it > 0) mbar_wait(&gfree_bar, (it - 1) & 1);   // every compute warp is done with the previous item's coefficients
        mbar_expect_tx(&g_bar, (unsigned)(sizeof(float) * GFLOATS));
        for (int w = 0; w < ND; ++w) {
          // mode 0 -> planes dy*ND.. at (y0, x0) ; mode 1 -> planes (2D-dy)*ND.. at (y0+dy-D, x0-D)
          const int plane0 = mode == 0 ? w * ND : (2 * D - w) * ND;
          tma_load_4d(gs + w * BOXG, &mapg, &g_bar, mode == 0 ? x0 : x0 - D, mode == 0 ? y0 : y0 + w - D, plane0, b);
        }
        const CUtensorMap* map = mode == 0 ? &map2 : &map1;  // the OTHER feature
        for (int gi = 0; gi < ngroups; ++gi, ++n) {
          const int s = n % RING;
          if (n >= RING) mbar_wait(&empty_bar[s], ((n / RING) - 1) & 1);
          mbar_expect_tx(&full_bar[s], (unsigned)(sizeof(float) * T::F2_STAGE));
          tma_load_4d(ring + s * T::F2_STAGE, map, &full_bar[s], x0 - D, y0 - D, gi * CC, b);
        }
      }
    }
    return;
  }

  // =========================== compute warps ===========================
  for (int it = 0; it < nit; ++it) {
    int b, mode, x0, y0;
    item_coord(it, b, mode, x0, y0);
    // LeakyReLU derivative: element (dy, dx, r, x) of the staged boxes is plane k at (sy, sx).  The sign bytes of this thread's
    // NE float4 groups are requested BEFORE the wait for the coefficient boxes (one L2 round trip, overlapped with the TMA
    // transfer); afterwards the staged copy is scaled in place (one block barrier per item).
    constexpr int G4 = GW / 4, NE = (ND * ND * TH * G4 + NW * 32 - 1) / (NW * 32);
    unsigned mword[(NE + 3) / 4];   // 4 sign bytes per register
#pragma unroll
    for (int j = 0; j < (NE + 3) / 4; ++j) mword[j] = 0u;
    if (mask != nullptr) {
      const int Wb = (W + 7) >> 3;
      const unsigned char* mp = mask + (size_t)b * ND * ND * H * Wb;
#pragma unroll
      for (int j = 0; j < NE; ++j) {
        const int e = tid + j * NW * 32;
        const int x4 = e % G4, r3 = e / G4;
        const int r = r3 % TH, pl = min(r3 / TH, ND * ND - 1);   // pl = dy * ND + dx (box order)
        const int dy = pl / ND, dxb = pl - dy * ND;
        const int k = mode == 0 ? pl : (2 * D - dy) * ND + dxb;
        const int sy = min(max(mode == 0 ? y0 + r : y0 + dy - D + r, 0), H - 1);
        const int sx = min(max((mode == 0 ? x0 : x0 - D) + 4 * x4, 0), W - 4);   // clamped: out-of-image coefficients are zero anyway
        mword[j >> 2] |= (unsigned)__ldg(mp + ((size_t)k * H + sy) * Wb + (sx >> 3)) << (8 * (j & 3));
      }
    }
    mbar_wait(&g_bar, it & 1);
    if (mask != nullptr) {
#pragma unroll
      for (int j = 0; j < NE; ++j) {
        const int e = tid + j * NW * 32;
        if (e < ND * ND * TH * G4) {
          const int x4 = e % G4;
          const int sx = min(max((mode == 0 ? x0 : x0 - D) + 4 * x4, 0), W - 4);
          const unsigned bits = ((mword[j >> 2] >> (8 * (j & 3))) & 0xffu) >> (sx & 7);
          float4* gp4 = reinterpret_cast<float4*>(gs + 4 * (size_t)e);   // e enumerates the float4 groups of the boxes in memory order
          float4 v = *gp4;
          if (!(bits & 1u)) v.x *= slope;
          if (!(bits & 2u)) v.y *= slope;
          if (!(bits & 4u)) v.z *= slope;
          if (!(bits & 8u)) v.w *= slope;
          *gp4 = v;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
    }
    float* dout = (mode == 0 ? df1 : df2) + (size_t)b * C * H * W;
    const int y = y0 + ty, xs = x0 + tx * PX;
    // flat stage index of this item's group 0 (every item has ngroups stages)
    const int n0 = it * ngroups;
    for (int gi = warp; gi < ngroups; gi += NW) {
      const int n = n0 + gi, s = n % RING;
      mbar_wait(&full_bar[s], (n / RING) & 1);
      const float* st = ring + s * T::F2_STAGE;
      float acc[CC][PX];
#pragma unroll
      for (int c = 0; c < CC; ++c)
#pragma unroll
        for (int p = 0; p < PX; ++p) acc[c][p] = 0.f;
#pragma unroll 1
      for (int dy = 0; dy < ND; ++dy) {
        float G[ND][PX];
        const float* gw = gs + dy * BOXG + ty * GW + tx * PX;
        if (mode == 0) {
#pragma unroll
          for (int dx = 0; dx < ND; ++dx)
#pragma unroll
            for (int q = 0; q < PX / 4; ++q) {
              const float4 v = *reinterpret_cast<const float4*>(gw + dx * TH * GW + 4 * q);
              G[dx][4 * q] = v.x; G[dx][4 * q + 1] = v.y; G[dx][4 * q + 2] = v.z; G[dx][4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
          for (int dx = 0; dx < ND; ++dx) {
            float t[PX + 4];
#pragma unroll
            for (int q = 0; q < PX / 4 + 1; ++q) {
              const float4 v = *reinterpret_cast<const float4*>(gw + (2 * D - dx) * TH * GW + (dx & ~3) + 4 * q);
              t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int p = 0; p < PX; ++p) G[dx][p] = t[p + (dx & 3)];
          }
        }
        const float* pw = st + (ty + dy) * S2 + tx * PX;
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          float w[WIN];
#pragma unroll
          for (int q = 0; q < WIN / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(pw + c * F2H * S2 + 4 * q);
            w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
          }
#pragma unroll
          for (int dx = 0; dx < ND; ++dx)
#pragma unroll
            for (int p = 0; p < PX; ++p) acc[c][p] = fmaf(G[dx][p], w[p + dx], acc[c][p]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);   // this warp was the only reader of the stage
      if (y < H && xs < W) {
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          const int ch = gi * CC + c;
          if (ch < C) {
            float* o = dout + ((size_t)ch * H + y) * W + xs;
#pragma unroll
            for (int q = 0; q < PX / 4; ++q)
              if (xs + 4 * q < W)
                *reinterpret_cast<float4*>(o + 4 * q) = make_float4(acc[c][4 * q] * inv_c, acc[c][4 * q + 1] * inv_c, acc[c][4 * q + 2] * inv_c, acc[c][4 * q + 3] * inv_c);
          }
        }
      }
    }
    // the coefficient area was modified with generic-proxy stores (mask pass): order them before the TMA (async proxy) refill
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(&gfree_bar);   // the coefficient area may be overwritten by the next item
  }
}

// generic-displacement backward: one thread per (b, c, y, x) element of d f1 / d f2.
__global__ void __launch_bounds__(128)
corr_bwd_generic(const float* __restrict__ g, const float* __restrict__ oact, const float* __restrict__ f1,
                 const float* __restrict__ f2, float* __restrict__ df1, float* __restrict__ df2, int C, int H, int W, int d,
                 long long g_bstride, long long a_bstride, float inv_c, float slope) {
  const int nd = 2 * d + 1;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= H * W) return;
  const int y = pix / W, x = pix - y * W;
  const int c = blockIdx.y, b = blockIdx.z;
  const size_t gb = (g_bstride ? (size_t)g_bstride : (size_t)nd * nd * H * W) * b;
  const size_t ab = (a_bstride ? (size_t)a_bstride : (size_t)nd * nd * H * W) * b;
  const size_t fb = ((size_t)b * C + c) * H * W;
  float a1 = 0.f, a2 = 0.f;
  for (int dyi = 0; dyi < nd; ++dyi) {
    const int dy = dyi - d;
    for (int dxi = 0; dxi < nd; ++dxi) {
      const int dx = dxi - d;
      const int k = dyi * nd + dxi;
      // d f1: g[k, y, x] * f2[y+dy, x+dx]
      const int yy = y + dy, xx = x + dx;
      if (df1 != nullptr && yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const size_t go = gb + ((size_t)k * H + y) * W + x;
        float gv = __ldg(g + go);
        if (oact != nullptr && !(__ldg(oact + (go - gb + ab)) > 0.f)) gv *= slope;
        a1 = fmaf(gv, __ldg(f2 + fb + (size_t)yy * W + xx), a1);
      }
      // d f2: g[k, y-dy, x-dx] * f1[y-dy, x-dx]
      const int ys = y - dy, xs = x - dx;
      if (df2 != nullptr && ys >= 0 && ys < H && xs >= 0 && xs < W) {
        const size_t go = gb + ((size_t)k * H + ys) * W + xs;
        float gv = __ldg(g + go);
        if (oact != nullptr && !(__ldg(oact + (go - gb + ab)) > 0.f)) gv *= slope;
        a2 = fmaf(gv, __ldg(f1 + fb + (size_t)ys * W + xs), a2);
      }
    }
  }
  if (df1 != nullptr) df1[fb + pix] = a1 * inv_c;
  if (df2 != nullptr) df2[fb + pix] = a2 * inv_c;
}

using Tile4 = CorrTile<4, 8, 4, 8, 8, 3>;
// d = 10: 4 x 32 pixel tile, thread = 4 pixels x 21 horizontal displacements (84 accumulators), 7 dy-warps per CTA, 3 CTAs per tile
using Tile10 = CorrTile<10, 4, 8, 4, 8, 3, 7, 128>;
using Tile4E = CorrTile<4, 8, 4, 8, 4, 6>;   // backward with the early first stage (GD = 3): 4-channel stages, 6 of them
#ifndef OCF_FWD_STAGES
#define OCF_FWD_STAGES 6
#endif
using Tile4P = CorrTile<4, 8, 4, 8, 8, OCF_FWD_STAGES>;  // persistent forward: one CTA per SM, deep ring
using Tile4PS = CorrTile<4, 8, 4, 8, 8, 4>;              // ... with the TMA-store epilogue: 4 stages + 81 staged output planes
constexpr int BWD_CR = 4;
constexpr int FWD_UNROLL = OCF_FWD_UNROLL;

// opt-in to > 48 KB of dynamic shared memory, once per kernel instantiation and size (the attribute is sticky; eager callers --
// the patched reference runs without a CUDA graph -- used to pay the driver call on every launch)
template <class K>
int set_smem(K kernel, size_t bytes) {
  static size_t done_for = 0;   // one instance per K (function-pointer type is not unique per kernel, hence the pointer check)
  static K done_kernel = nullptr;
  if (bytes > 48 * 1024 && !(done_kernel == kernel && done_for == bytes)) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    done_kernel = kernel;
    done_for = bytes;
  }
  return 0;
}

// Channel split factor (power of two <= 8), from a rounds-of-work cost model in units of "one channel of one tile":
//   cost(ks) = ceil(tiles * ks / slots) * (prologue + C / ks) + (ks > 1 ? reduce : 0)
// `prologue` is the fixed per-CTA cost expressed in channels.  Forward: the first TMA box (~4 channels) and, when split,
// the DSMEM reduction of the 81 partial planes (~0.6 of a full tile).  Backward: staging + lifting the 81 coefficient
// planes of the tile (and the LeakyReLU mask planes) costs as much as ~30 (16 without the mask) channels of the main
// loop (measured: t(C) = 11 us + 0.36 us * C per round at the L2 geometry, tools/probe_corr.py), so splitting the
// channels of an already full grid only multiplies that prologue.  At least `min_ch` channels per CTA.
int pick_ksplit(long long tiles, int C, int min_ch, bool fwd, double prologue) {
  if (const char* e = getenv(fwd ? "OCF_KSPLIT_FWD" : "OCF_KSPLIT_BWD")) {  // developer override for tuning runs
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) return v;
  }
  const long long slots = 2LL * OCF_SM_COUNT;  // 2 resident CTAs per SM
  int best = 1;
  double best_cost = 1e30;
  for (int ks = 1; ks <= 8; ks *= 2) {
    if (ks > 1 && C / ks < min_ch) break;
    // measured (brute-force sweep over ks at every pyramid level): 8-wide clusters stop paying once there are more than
    // ~100 CTAs of them -- L5 (16 tiles): ks = 8 -> 128 CTAs 22.5 us, ks = 4 -> 64 CTAs 18.4 us; L6 (8 tiles): ks = 8 -> 64 CTAs
    // 18.4 us, ks = 4 -> 24.6 us -- so the forward counts only 96 placement slots for them
    const long long eff_slots = (fwd && ks == 8) ? 96 : slots;
    const long long rounds = (tiles * ks + eff_slots - 1) / eff_slots;
    const double per_cta = prologue + (double)((C + ks - 1) / ks);
    const double cost = (double)rounds * per_cta + ((ks > 1 && fwd) ? 0.6 * C : 0.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = ks; }
  }
  return best;
}

template <class K, class... Args>
int launch_kernel(K kernel, dim3 grid, int threads, size_t smem, cudaStream_t s, int cluster_z, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = cluster_z;
  cfg.attrs = attr;
  cfg.numAttrs = cluster_z > 1 ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  return e == cudaSuccess ? 0 : (int)e;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (the .so must load on a machine without libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// NCHW fp32 tensor [B, C, H, W] -> 4-D tensor map with a {bw, bh, cc, 1} box; out-of-bounds elements read as zero
// Descriptors are pure functions of (pointer, shape, box, stride): the last few are kept per host thread, so a training loop
// that re-presents the same activations (torch's caching allocator hands back the same blocks every step) encodes each map once.
struct MapKey {
  const float* base;
  int B, C, H, W, bw, bh, cc;
  long long bstride;
  bool operator==(const MapKey& o) const {
    return base == o.base && B == o.B && C == o.C && H == o.H && W == o.W && bw == o.bw && bh == o.bh && cc == o.cc && bstride == o.bstride;
  }
};
constexpr int MAP_CACHE = 64;
struct MapCache {
  MapKey key[MAP_CACHE];
  CUtensorMap map[MAP_CACHE];
  int used = 0, next = 0;
};

bool make_map_uncached(CUtensorMap* map, const float* base, int B, int C, int H, int W, int bw, int bh, int cc, long long bstride);

bool make_map(CUtensorMap* map, const float* base, int B, int C, int H, int W, int bw, int bh, int cc, long long bstride = 0) {
  static thread_local MapCache cache;
  const MapKey k{base, B, C, H, W, bw, bh, cc, bstride};
  for (int i = 0; i < cache.used; ++i)
    if (cache.key[i] == k) { *map = cache.map[i]; return true; }
  if (!make_map_uncached(map, base, B, C, H, W, bw, bh, cc, bstride)) return false;
  const int slot = cache.next;
  cache.key[slot] = k;
  cache.map[slot] = *map;
  cache.next = (slot + 1) % MAP_CACHE;
  if (cache.used < MAP_CACHE) ++cache.used;
  return true;
}

bool make_map_uncached(CUtensorMap* map, const float* base, int B, int C, int H, int W, int bw, int bh, int cc, long long bstride) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)(bstride ? bstride : (long long)W * H * C) * 4};
  const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)cc, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool ocf_make_tensor_map(CUtensorMap* map, const float* base, int B, int C, int H, int W, int bw, int bh, int cc, long long bstride) {
  return make_map(map, base, B, C, H, W, bw, bh, cc, bstride);
}

// f1_bstride: batch stride of f1 in elements (0 = dense) -- the fused level op keeps the normalised first feature map inside
// the decoder's concat buffer (ops.level_fused); only the TMA path (d = 4, 16-byte aligned rows) reads it in place
static int corr_fwd_impl(const float* f1, long long f1_bstride, const float* f2, float* out, int B, int C, int H, int W, int d,
                         long long out_bstride, float leaky_slope, const float* norm, unsigned char* mask_out,
                         ocf_stream_t stream) {
  OCF_REQUIRE_PTR(f1); OCF_REQUIRE_PTR(f2); OCF_REQUIRE_PTR(out);
  OCF_REQUIRE(f1_bstride == 0 || f1_bstride >= (long long)C * H * W, OCF_ESHAPE);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(d >= 0 && d <= OCF_MAX_DISPLACEMENT, OCF_EUNSUPPORTED);
  const long long nd = 2 * d + 1;
  OCF_REQUIRE(out_bstride == 0 || out_bstride >= nd * nd * H * W, OCF_ESHAPE);
  OCF_REQUIRE(B <= 8191, OCF_EUNSUPPORTED);
  OCF_REQUIRE(mask_out == nullptr || d == 4, OCF_EUNSUPPORTED);  // only the d = 4 kernels emit the sign mask
  cudaStream_t s = ocf_cast_stream(stream);
  const float inv_c = 1.0f / (float)C;
  // d = 4 has two implementations: the fp32 FMA kernels below (TMA-fed; 1e-7 relative error) and a tensor-core kernel
  // (corr_tc.cu: tcgen05, 3xTF32, 1e-6).  Measured on B200 the FMA kernels win wherever TMA can describe the rows (W % 4 == 0:
  // 29.7 vs 49.9 us at 8x32x96x128, 79 vs 133 us at C = 128) and lose on ragged rows, where they fall back to 4-byte cp.async
  // (8x32x188x621: 660 vs 456 us) -- so ragged rows take the tensor-core kernel.  OCF_CORR_TC=1 forces it everywhere (tuning).
  static const int use_tc = []() { const char* e = getenv("OCF_CORR_TC"); return e ? atoi(e) : 0; }();
  if (d == 4 && (use_tc || W % 4 != 0) && f1_bstride == 0 && norm == nullptr) return ocf_corr_fwd_tc_launch(f1, f2, out, mask_out, norm, nullptr, 0, nullptr, B, C, H, W, out_bstride, leaky_slope, s);
  if (d == 4 && norm == nullptr) {
    using T = Tile4;
    const size_t smem = sizeof(float) * T::STAGES * (T::F1_STAGE + T::F2_STAGE);
    const int gx = (W + T::TW - 1) / T::TW, gy = (H + T::TH - 1) / T::TH;
    // one tile or more per SM: the persistent kernel (ks == 1) keeps every SM busy without any reduction
    const int ks = (long long)gx * gy * B >= OCF_SM_COUNT ? 1 : pick_ksplit((long long)gx * gy * B, C, 2 * T::CC, true, 4.0);
    dim3 grid(gx, gy, B * ks);
    const bool vec = (W % 4 == 0) && ocf_aligned16(f1) && ocf_aligned16(f2) && ocf_aligned16(out) && (out_bstride % 4 == 0);
    CUtensorMap m1, m2;
    memset(&m1, 0, sizeof(m1));
    memset(&m2, 0, sizeof(m2));
    const bool tma = vec && (f1_bstride % 4 == 0) && make_map(&m1, f1, B, C, H, W, T::S1, T::TH, T::CC, f1_bstride) &&
                     make_map(&m2, f2, B, C, H, W, T::S2, T::F2H, T::CC);
    OCF_REQUIRE(tma || f1_bstride == 0, OCF_EUNSUPPORTED);
    static const int tma_store = []() { const char* e = getenv("OCF_FWD_TMA_STORE"); return e ? atoi(e) : 1; }();   // developer knob (A/B runs)
    CUtensorMap mo;
    memset(&mo, 0, sizeof(mo));
    if (tma && ks == 1 && tma_store && make_map(&mo, out, B, 81, H, W, T::TW, T::TH, T::ND, out_bstride)) {
      using TP = Tile4PS;
      auto kernel = corr_fwd_persist<TP, FWD_UNROLL, true>;
      const size_t psmem = sizeof(float) * (TP::STAGES * (TP::F1_STAGE + TP::F2_STAGE) + TP::ND * TP::ND * TP::TH * TP::TW);
      if (int e = set_smem(kernel, psmem)) return e;
      const int ntiles = gx * gy * B;
      const int nblk = ntiles < OCF_FWD_CTAS_PER_SM * OCF_SM_COUNT ? ntiles : OCF_FWD_CTAS_PER_SM * OCF_SM_COUNT;
      if (int e = launch_kernel(kernel, dim3(nblk), TP::THREADS + 32, psmem, s, 1, m1, m2, mo, out, mask_out, C, H, W, out_bstride, inv_c, leaky_slope, gx, gy, ntiles)) return e;
    } else if (tma && ks == 1) {
      using TP = Tile4P;
      auto kernel = corr_fwd_persist<TP, FWD_UNROLL>;
      const size_t psmem = sizeof(float) * TP::STAGES * (TP::F1_STAGE + TP::F2_STAGE);
      if (int e = set_smem(kernel, psmem)) return e;
      const int ntiles = gx * gy * B;
      const int nblk = ntiles < OCF_FWD_CTAS_PER_SM * OCF_SM_COUNT ? ntiles : OCF_FWD_CTAS_PER_SM * OCF_SM_COUNT;
      if (int e = launch_kernel(kernel, dim3(nblk), TP::THREADS + 32, psmem, s, 1, m1, m2, mo, out, mask_out, C, H, W, out_bstride, inv_c, leaky_slope, gx, gy, ntiles)) return e;
    } else if (tma) {
      auto kernel = corr_fwd_tiled<T, STG_TMA>;
      if (int e = set_smem(kernel, smem)) return e;
      if (int e = launch_kernel(kernel, grid, T::THREADS, smem, s, ks, m1, m2, f1, f2, out, mask_out, C, H, W, out_bstride, inv_c, leaky_slope, ks)) return e;
    } else {
      auto kernel = vec ? corr_fwd_tiled<T, STG_ASYNC16> : corr_fwd_tiled<T, STG_ASYNC4>;
      if (int e = set_smem(kernel, smem)) return e;
      if (int e = launch_kernel(kernel, grid, T::THREADS, smem, s, ks, m1, m2, f1, f2, out, mask_out, C, H, W, out_bstride, inv_c, leaky_slope, ks)) return e;
    }
  } else if (d == 10 && norm == nullptr && f1_bstride == 0) {
    // FlowNetC family (flow_net_c.py:22-25, flow_occ_net_c.py:26, occlusion_net_c.py:24): 21 x 21 = 441 planes
    using T = Tile10;
    const size_t smem = sizeof(float) * T::STAGES * (T::F1_STAGE + T::F2_STAGE);
    const int gx = (W + T::TW - 1) / T::TW, gy = (H + T::TH - 1) / T::TH;
    OCF_REQUIRE((long long)B * T::NGY <= 65535, OCF_EUNSUPPORTED);
    dim3 grid(gx, gy, B * T::NGY);
    const bool vec = (W % 4 == 0) && ocf_aligned16(f1) && ocf_aligned16(f2) && ocf_aligned16(out) && (out_bstride % 4 == 0);
    CUtensorMap m1, m2;
    memset(&m1, 0, sizeof(m1));
    memset(&m2, 0, sizeof(m2));
    // The TMA-staged instantiation of this tile FAULTS on B200 ("illegal instruction" at the first box; the d = 4 instantiation of the
    // same kernel is fine and the cp.async instantiations of this tile are correct), cause not found yet: it stays behind a
    // developer knob and d = 10 is staged with 16-byte (f1 tile) and 8-byte (halo, which starts at x0 - 10) cp.async copies.
    static const int use_tma = []() { const char* e = getenv("OCF_CORR10_TMA"); return e ? atoi(e) : 0; }();
    const bool tma = vec && use_tma && make_map(&m1, f1, B, C, H, W, T::S1, T::TH, T::CC) && make_map(&m2, f2, B, C, H, W, T::S2, T::F2H, T::CC);
    auto kernel = tma ? corr_fwd_tiled<T, STG_TMA> : (vec ? corr_fwd_tiled<T, STG_ASYNC16> : corr_fwd_tiled<T, STG_ASYNC4>);
    if (int e = set_smem(kernel, smem)) return e;
    if (int e = launch_kernel(kernel, grid, T::THREADS, smem, s, 1, m1, m2, f1, f2, out, (unsigned char*)nullptr, C, H, W, out_bstride, inv_c, leaky_slope, 1)) return e;
  } else {
    OCF_REQUIRE(f1_bstride == 0, OCF_EUNSUPPORTED);
    dim3 grid((H * W + 127) / 128, (unsigned)nd, B);
    corr_fwd_generic<<<grid, 128, 0, s>>>(f1, f2, out, C, H, W, d, out_bstride, inv_c, leaky_slope, norm);
  }
  return ocf_launch_status();
}

extern "C" int ocf_corr_fwd(const float* f1, const float* f2, float* out, int B, int C, int H, int W, int d,
                            long long out_bstride, float leaky_slope, const float* norm, unsigned char* mask_out,
                            ocf_stream_t stream) {
  return corr_fwd_impl(f1, 0, f2, out, B, C, H, W, d, out_bstride, leaky_slope, norm, mask_out, stream);
}

extern "C" int ocf_corr_fwd_strided(const float* f1, long long f1_bstride, const float* f2, float* out, long long out_bstride,
                                    unsigned char* mask_out, int B, int C, int H, int W, float leaky_slope, ocf_stream_t stream) {
  OCF_REQUIRE(W % 4 == 0, OCF_EUNSUPPORTED);
  return corr_fwd_impl(f1, f1_bstride, f2, out, B, C, H, W, 4, out_bstride, leaky_slope, nullptr, mask_out, stream);
}

// f1_bstride / f2_bstride: batch strides of the feature operands in elements (0 = dense): the fused level op keeps the
// normalised first feature map inside the decoder's concat buffer (level.cu)
int ocf_corr_bwd_impl(const float* grad_out, const float* out_act, const float* f1, const float* f2, float* df1,
                      float* df2, int B, int C, int H, int W, int d, long long g_bstride, long long act_bstride,
                      float leaky_slope, const unsigned char* mask, long long f1_bstride, long long f2_bstride, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(grad_out); OCF_REQUIRE_PTR(f1); OCF_REQUIRE_PTR(f2);
  OCF_REQUIRE(df1 != nullptr || df2 != nullptr, OCF_ENULL);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(d >= 0 && d <= OCF_MAX_DISPLACEMENT, OCF_EUNSUPPORTED);
  const long long nd = 2 * d + 1;
  OCF_REQUIRE(g_bstride == 0 || g_bstride >= nd * nd * H * W, OCF_ESHAPE);
  OCF_REQUIRE(act_bstride == 0 || act_bstride >= nd * nd * H * W, OCF_ESHAPE);
  OCF_REQUIRE(B <= 4095 && C <= 65535, OCF_EUNSUPPORTED);
  OCF_REQUIRE(mask == nullptr || d == 4, OCF_EUNSUPPORTED);
  if (mask != nullptr) out_act = nullptr;  // the bitmask carries the same information in 1/32 of the bytes
  cudaStream_t s = ocf_cast_stream(stream);
  const float inv_c = 1.0f / (float)C;
  if (d == 4) {
    using T = Tile4;
    size_t smem = sizeof(float) * (T::STAGES * T::F2_STAGE + T::ND * BWD_CR * T::TH * T::S1);
    const int nmodes = (df1 != nullptr && df2 != nullptr) ? 2 : 1;
    const int first = df1 != nullptr ? 0 : 1;
    const int gx = (W + T::TW - 1) / T::TW, gy = (H + T::TH - 1) / T::TH;
    const int ks = pick_ksplit((long long)gx * gy * B * nmodes, C, 2 * T::CC, false, out_act != nullptr ? 30.0 : (mask != nullptr ? 18.0 : 16.0));
    dim3 grid(gx, gy, B * nmodes * ks);
    const bool vec = (W % 4 == 0) && (f1_bstride % 4 == 0) && (f2_bstride % 4 == 0) && ocf_aligned16(f1) && ocf_aligned16(f2) && (df1 == nullptr || ocf_aligned16(df1)) &&
                     (df2 == nullptr || ocf_aligned16(df2));
    CUtensorMap m1, m2, mg;
    memset(&m1, 0, sizeof(m1));
    memset(&m2, 0, sizeof(m2));
    memset(&mg, 0, sizeof(mg));
    const long long gbs = g_bstride ? g_bstride : nd * nd * H * W;
    const long long abs_ = act_bstride ? act_bstride : nd * nd * H * W;
    bool tma = vec && ocf_aligned16(grad_out) && (out_act == nullptr || ocf_aligned16(out_act)) && (gbs % 4 == 0) && (abs_ % 4 == 0);
    tma = tma && (f1_bstride % 4 == 0) && (f2_bstride % 4 == 0);
    tma = tma && (long long)C * H * W < (1LL << 31);   // 32-bit element offsets inside one batch item (the reduction's stores)
    tma = tma && make_map(&m1, f1, B, C, H, W, T::S2, T::F2H, T::CC, f1_bstride) && make_map(&m2, f2, B, C, H, W, T::S2, T::F2H, T::CC, f2_bstride) &&
          make_map(&mg, grad_out, B, (int)(nd * nd), H, W, T::S2, T::TH, T::ND, gbs);
    // developer knob (A/B runs): OCF_BWD_GDIRECT=1 loads the coefficients global -> registers.  Measured: SLOWER (L2 level 70.3 vs
    // 61.3 us, L3 43.0 vs 38.8): with 96 registers per thread only ~5 of the 18-24 128-bit loads are in flight, so the lift costs
    // several L2 round trips instead of one TMA round trip.  The staged form stays the default.
    // OCF_BWD_GDIRECT=2: direct for mode 0 only; =3: early first stage (4-channel stages, 40-float boxes).  OCF_BWD_PREFETCH (default 1):
    // L2 prefetch of the next wave's boxes (stride = the 2 x 148 resident CTAs of a wave; any other positive value is taken as the stride
    // itself; 0 = off).  OCF_KNOBS_DYNAMIC=1 re-reads both on every call (A/B probes inside one process: tools/probe_bwd_knobs.py);
    // otherwise they are read once.  Measured with the per-CTA timeline of tools/timeline.py (profiles/r2_corr_bwd_timeline.txt; L2
    // level 8x32x96x128, times per CTA): 17.4-17.9 us = 3.0-4.2 us until the coefficient boxes have landed + 0.6 lift + 1.2-1.4 first
    // feature stage + 10-10.5 main loop (7.6 when the CTA is alone on its SM); three generations of CTAs per slot.
    //   prefetch:      later-wave boxes land after 2.1 instead of 3.0 us, first stage after 0.7 instead of 1.4: kernel span 55.9 -> 53.6 us,
    //                  call 61.4 -> 59.4 us (8x32x188x620: 455.7 -> 447.5 us); bit-identical; DEFAULT.
    //   GDIRECT=2:     65.6 us -- the direct lift of mode 0 takes 4.2-5.8 us (TMA + lift: 3.6-4.8) and this instantiation's main loop 13.3 us.
    //   GDIRECT=3:     60.4 us, with prefetch 59.4 (= prefetch alone): the first stage is there 1.1-1.3 us earlier, but the 4-channel stages
    //                  cost the main loop 0.4-0.8 us and the lift has to wait for all nine warps (the reduction buffer aliases the boxes).
    static const bool knobs_dynamic = getenv("OCF_KNOBS_DYNAMIC") != nullptr;
    auto knob = [](const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; };
    auto knob1 = [](const char* name) { const char* e = getenv(name); return e ? atoi(e) : 1; };   // default ON
    static const int gdirect0 = knob("OCF_BWD_GDIRECT"), prefetch0 = knob1("OCF_BWD_PREFETCH");
    const int gdirect = knobs_dynamic ? knob("OCF_BWD_GDIRECT") : gdirect0;
    const int prefetch = knobs_dynamic ? knob1("OCF_BWD_PREFETCH") : prefetch0;
    const int pf_stride = prefetch <= 0 ? 0 : (prefetch == 1 ? 2 * OCF_SM_COUNT : prefetch);
    // channel-split persistent form (experiment, OCF_BWD_CS=1): full grids only, sign bitmask or no activation.  Measured SLOWER
    // than the tiled form (L2 level 79 vs 60 us, C = 64: 132 vs 101, C = 128: 228 vs 181): with the coefficients resident in
    // shared memory only 4 feature stages fit, i.e. 4 compute warps per SM -- one per scheduler, issue_active 40 %, fma 29 % -- and
    // every item pays a serial coefficient refill.  No barriers in the main loop, but too little parallelism to profit.
    static const int use_cs = []() { const char* e = getenv("OCF_BWD_CS"); return e ? atoi(e) : 0; }();
    const long long nitems = (long long)gx * gy * B * nmodes;
    if (tma && use_cs && out_act == nullptr && nitems >= OCF_SM_COUNT && nitems < (1LL << 30)) {
      constexpr int NW = 4, RING = 4;
      auto kernel = corr_bwd_cs<T, NW, RING>;
      const size_t csmem = sizeof(float) * ((size_t)T::ND * T::ND * T::TH * T::S2 + (size_t)RING * T::F2_STAGE);
      if (int e = set_smem(kernel, csmem)) return e;
      if (int e = launch_kernel(kernel, dim3(OCF_SM_COUNT), (NW + 1) * 32, csmem, s, 1, m1, m2, mg, mask, df1, df2, C, H, W, inv_c, leaky_slope, nmodes,
                                first, gx, gy, (int)nitems)) return e;
    } else if (tma && gdirect == 3) {
      // early first stage: own tile type (4-channel stages) and maps (40-float coefficient boxes, 4-channel feature boxes)
      using E = Tile4E;
      CUtensorMap e1, e2, eg;
      if (!(make_map(&e1, f1, B, C, H, W, E::S2, E::F2H, E::CC, f1_bstride) && make_map(&e2, f2, B, C, H, W, E::S2, E::F2H, E::CC, f2_bstride) &&
            make_map(&eg, grad_out, B, (int)(nd * nd), H, W, E::F2W, E::TH, E::ND, gbs))) return OCF_EUNSUPPORTED;
      const size_t esmem = sizeof(float) * ((size_t)E::ND * E::ND * E::TH * E::F2W + E::F2_STAGE);
      static_assert(sizeof(float) * ((E::STAGES - 1) * E::F2_STAGE + E::ND * BWD_CR * E::TH * E::S1) <= sizeof(float) * ((size_t)E::ND * E::ND * E::TH * E::F2W),
                    "ring stages 1.. and the reduction buffer must fit inside the staging area");
      auto kernel = corr_bwd_tiled<E, BWD_CR, STG_TMA, 3>;
      if (int e = set_smem(kernel, esmem)) return e;
      if (int e = launch_kernel(kernel, grid, E::THREADS + 32, esmem, s, 1, e1, e2, eg, grad_out, out_act, mask, f1, f2, df1, df2, C, H, W, g_bstride,
                                act_bstride, inv_c, leaky_slope, nmodes, first, ks, f1_bstride, f2_bstride, pf_stride)) return e;
    } else if (tma && gdirect == 1) {
      auto kernel = corr_bwd_tiled<T, BWD_CR, STG_TMA, 1>;
      if (int e = set_smem(kernel, smem)) return e;
      if (int e = launch_kernel(kernel, grid, T::THREADS + 32, smem, s, 1, m1, m2, mg, grad_out, out_act, mask, f1, f2, df1, df2, C, H, W, g_bstride,
                                act_bstride, inv_c, leaky_slope, nmodes, first, ks, f1_bstride, f2_bstride, pf_stride)) return e;
    } else if (tma) {
      const size_t gstage = sizeof(float) * T::ND * T::ND * T::TH * T::S2;  // coefficient staging, reused by ring + red
      if (gstage > smem) smem = gstage;
      auto kernel = gdirect == 2 ? corr_bwd_tiled<T, BWD_CR, STG_TMA, 2> : corr_bwd_tiled<T, BWD_CR, STG_TMA>;
      if (int e = set_smem(kernel, smem)) return e;
      if (int e = launch_kernel(kernel, grid, T::THREADS + 32, smem, s, 1, m1, m2, mg, grad_out, out_act, mask, f1, f2, df1, df2, C, H, W, g_bstride,
                                act_bstride, inv_c, leaky_slope, nmodes, first, ks, f1_bstride, f2_bstride, pf_stride)) return e;
    } else {
      auto kernel = vec ? corr_bwd_tiled<T, BWD_CR, STG_ASYNC16> : corr_bwd_tiled<T, BWD_CR, STG_ASYNC4>;
      if (int e = set_smem(kernel, smem)) return e;
      if (int e = launch_kernel(kernel, grid, T::THREADS, smem, s, 1, m1, m2, mg, grad_out, out_act, mask, f1, f2, df1, df2, C, H, W, g_bstride,
                                act_bstride, inv_c, leaky_slope, nmodes, first, ks, f1_bstride, f2_bstride, 0)) return e;
    }
  } else if (d == 10 && f1_bstride == 0 && f2_bstride == 0) {
    // FlowNetC family: three dy-group CTAs per tile and mode, partial sums accumulated with vector reds into the zeroed outputs
    using T = Tile10;
    const size_t smem = sizeof(float) * (T::STAGES * T::F2_STAGE + T::NDY * BWD_CR * T::TH * T::S1);
    const int nmodes = (df1 != nullptr && df2 != nullptr) ? 2 : 1;
    const int first = df1 != nullptr ? 0 : 1;
    const int gx = (W + T::TW - 1) / T::TW, gy = (H + T::TH - 1) / T::TH;
    OCF_REQUIRE((long long)B * nmodes * T::NGY <= 65535, OCF_EUNSUPPORTED);
    dim3 grid(gx, gy, B * nmodes * T::NGY);
    const bool vec = (W % 4 == 0) && ocf_aligned16(f1) && ocf_aligned16(f2) && (df1 == nullptr || ocf_aligned16(df1)) && (df2 == nullptr || ocf_aligned16(df2));
    cudaError_t e;
    if (df1 != nullptr && (e = cudaMemsetAsync(df1, 0, sizeof(float) * (size_t)B * C * H * W, s)) != cudaSuccess) return (int)e;
    if (df2 != nullptr && (e = cudaMemsetAsync(df2, 0, sizeof(float) * (size_t)B * C * H * W, s)) != cudaSuccess) return (int)e;
    CUtensorMap m1, m2, mg;
    memset(&m1, 0, sizeof(m1));
    memset(&m2, 0, sizeof(m2));
    memset(&mg, 0, sizeof(mg));
    auto kernel = vec ? corr_bwd_tiled<T, BWD_CR, STG_ASYNC16> : corr_bwd_tiled<T, BWD_CR, STG_ASYNC4>;
    if (int er = set_smem(kernel, smem)) return er;
    if (int er = launch_kernel(kernel, grid, T::THREADS, smem, s, 1, m1, m2, mg, grad_out, out_act, mask, f1, f2, df1, df2, C, H, W, g_bstride,
                               act_bstride, inv_c, leaky_slope, nmodes, first, 1, f1_bstride, f2_bstride, 0)) return er;
  } else {
    OCF_REQUIRE(f1_bstride == 0 && f2_bstride == 0, OCF_EUNSUPPORTED);
    dim3 grid((H * W + 127) / 128, C, B);
    corr_bwd_generic<<<grid, 128, 0, s>>>(grad_out, out_act, f1, f2, df1, df2, C, H, W, d, g_bstride, act_bstride, inv_c, leaky_slope);
  }
  return ocf_launch_status();
}

extern "C" int ocf_corr_bwd(const float* grad_out, const float* out_act, const float* f1, const float* f2, float* df1,
                            float* df2, int B, int C, int H, int W, int d, long long g_bstride, long long act_bstride,
                            float leaky_slope, const unsigned char* mask, ocf_stream_t stream) {
  return ocf_corr_bwd_impl(grad_out, out_act, f1, f2, df1, df2, B, C, H, W, d, g_bstride, act_bstride, leaky_slope, mask, 0, 0, stream);
}
