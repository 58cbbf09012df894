// Local cost volume (PWC-style correlation), forward and backward, fp32, NCHW.
//
// Replaces compute_cost_volume (reference models/networks/correlation_layer.py:7-40): 81 x
// (slice, mul, mean) + cat => one launch.  See DESIGN.md "corr" for the tiling rationale.
//
//   forward : a CTA owns a TH x TW pixel tile of one batch item.  Per channel chunk it stages the f1
//             tile and the f2 tile + halo in shared memory (row strides == 4 mod 8 floats so that the
//             128-bit window loads of a quarter-warp hit 8 distinct bank groups).  Warp w of the CTA
//             owns the vertical displacement dy = w - d; a thread owns PX consecutive pixels of one
//             row and all 2d+1 horizontal displacements: (2d+1)*PX register accumulators,
//             (PX + PX+2d)/4 LDS.128 per (2d+1)*PX FMAs.
//   backward: d f1 and d f2 are both written as gathers (deterministic, no atomics):
//               d f1[c,p] = 1/C sum_delta g[delta, p]        * f2[c, p+delta]
//               d f2[c,q] = 1/C sum_delta g[-delta, q+delta] * f1[c, q+delta]
//             i.e. the same "window of the other feature times 81 per-pixel coefficients"; only the
//             way the coefficients are fetched differs.  A thread keeps its (2d+1)*PX coefficients
//             in registers for the whole kernel, the 2d+1 dy-warps reduce through shared memory.
#include "common.cuh"

namespace {

constexpr int pad4mod8(int n) {
  int r = (n + 3) / 4 * 4;
  return (r % 8 == 4) ? r : r + 4;
}

template <int D_, int PX_, int TXT_, int TH_, int CC_>
struct CorrTile {
  static constexpr int D = D_, PX = PX_, TXT = TXT_, TH = TH_, CC = CC_;
  static constexpr int ND = 2 * D + 1;
  static constexpr int TW = PX * TXT;
  static constexpr int F2W = TW + 2 * D;
  static constexpr int F2H = TH + 2 * D;
  static constexpr int S1 = pad4mod8(TW);
  static constexpr int S2 = pad4mod8(F2W);
  static constexpr int LANES = TXT * TH;  // pixel-threads per dy (one warp when == 32)
  static constexpr int THREADS = LANES * ND;
  static constexpr int WIN = PX + 2 * D;  // f2 window per thread
  static_assert(PX % 4 == 0 && (2 * D) % 4 == 0, "128-bit window loads need PX and 2D multiples of 4");
  static_assert(LANES == 32, "one warp per vertical displacement");
};

// ---- tile staging -------------------------------------------------------------------------------
// Copies a [CC][ROWS][COLS] box of a NCHW tensor (origin (yb, xb), may be negative / past the edge)
// into shared memory with row stride S, zero-filling everything outside the image or past channel C.
// VEC: rows are 16-byte aligned in global memory (W % 4 == 0, xb % 4 == 0, base aligned).
template <int CC, int ROWS, int COLS, int S, int THREADS, bool VEC>
__device__ __forceinline__ void stage_box(float* __restrict__ dst, const float* __restrict__ src_b, int c0, int C,
                                          int H, int W, int yb, int xb, float nmean, float ninv, bool do_norm) {
  if (VEC) {
    constexpr int C4 = COLS / 4;
    constexpr int ITEMS = CC * ROWS * C4;
    for (int i = threadIdx.x; i < ITEMS; i += THREADS) {
      const int row = i / C4, q = i - row * C4;
      const int c = row / ROWS, r = row - c * ROWS;
      const int gy = yb + r, gx = xb + q * 4, gc = c0 + c;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gc < C && gy >= 0 && gy < H && gx >= 0 && gx < W) {
        v = __ldg(reinterpret_cast<const float4*>(src_b + ((size_t)gc * H + gy) * W + gx));
        if (do_norm) {
          v.x = (v.x - nmean) * ninv; v.y = (v.y - nmean) * ninv;
          v.z = (v.z - nmean) * ninv; v.w = (v.w - nmean) * ninv;
        }
      }
      *reinterpret_cast<float4*>(dst + (c * ROWS + r) * S + q * 4) = v;
    }
  } else {
    constexpr int ITEMS = CC * ROWS * COLS;
    for (int i = threadIdx.x; i < ITEMS; i += THREADS) {
      const int row = i / COLS, x = i - row * COLS;
      const int c = row / ROWS, r = row - c * ROWS;
      const int gy = yb + r, gx = xb + x, gc = c0 + c;
      float v = 0.f;
      if (gc < C && gy >= 0 && gy < H && gx >= 0 && gx < W) {
        v = __ldg(src_b + ((size_t)gc * H + gy) * W + gx);
        if (do_norm) v = (v - nmean) * ninv;
      }
      dst[(c * ROWS + r) * S + x] = v;
    }
  }
}

// ---- forward ------------------------------------------------------------------------------------
template <class T, bool VEC>
__global__ void __launch_bounds__(T::THREADS, 2)
corr_fwd_tiled(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ out, int C, int H, int W,
               long long out_bstride, float inv_c, float slope, const float* __restrict__ norm) {
  constexpr int D = T::D, PX = T::PX, ND = T::ND, TH = T::TH, TW = T::TW, CC = T::CC;
  constexpr int S1 = T::S1, S2 = T::S2, F2H = T::F2H, F2W = T::F2W, WIN = T::WIN;
  extern __shared__ __align__(16) float smem[];
  float* f1s = smem;                 // [CC][TH][S1]
  float* f2s = smem + CC * TH * S1;  // [CC][F2H][S2]

  const int tid = threadIdx.x;
  const int lane = tid % T::LANES;
  const int tx = lane % T::TXT, ty = lane / T::TXT;
  const int dyi = tid / T::LANES;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
  const float* f1b = f1 + (size_t)b * C * H * W;
  const float* f2b = f2 + (size_t)b * C * H * W;
  float nmean = 0.f, ninv = 1.f;
  const bool do_norm = norm != nullptr;
  if (do_norm) { nmean = __ldg(norm); ninv = __ldg(norm + 1); }

  float acc[ND][PX];
#pragma unroll
  for (int i = 0; i < ND; ++i)
#pragma unroll
    for (int p = 0; p < PX; ++p) acc[i][p] = 0.f;

  for (int c0 = 0; c0 < C; c0 += CC) {
    __syncthreads();
    stage_box<CC, TH, TW, S1, T::THREADS, VEC>(f1s, f1b, c0, C, H, W, y0, x0, nmean, ninv, do_norm);
    stage_box<CC, F2H, F2W, S2, T::THREADS, VEC>(f2s, f2b, c0, C, H, W, y0 - D, x0 - D, nmean, ninv, do_norm);
    __syncthreads();
    const float* p1 = f1s + ty * S1 + tx * PX;
    const float* p2 = f2s + (ty + dyi) * S2 + tx * PX;
#pragma unroll 2
    for (int c = 0; c < CC; ++c) {
      float a[PX], w[WIN];
#pragma unroll
      for (int i = 0; i < PX / 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(p1 + c * TH * S1 + 4 * i);
        a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < WIN / 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(p2 + c * F2H * S2 + 4 * i);
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
      }
#pragma unroll
      for (int dx = 0; dx < ND; ++dx)
#pragma unroll
        for (int p = 0; p < PX; ++p) acc[dx][p] = fmaf(a[p], w[p + dx], acc[dx][p]);
    }
  }

  const int y = y0 + ty, xs = x0 + tx * PX;
  if (y >= H || xs >= W) return;
  const size_t bstride = out_bstride ? (size_t)out_bstride : (size_t)ND * ND * H * W;
  float* ob = out + (size_t)b * bstride + ((size_t)(dyi * ND) * H + y) * W + xs;
#pragma unroll
  for (int dx = 0; dx < ND; ++dx) {
    float r[PX];
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      float v = acc[dx][p] * inv_c;
      r[p] = v > 0.f ? v : v * slope;
    }
    float* o = ob + (size_t)dx * H * W;
    if (VEC) {  // W % 4 == 0 and 16B-aligned rows: each float4 is entirely inside or outside
#pragma unroll
      for (int i = 0; i < PX / 4; ++i)
        if (xs + 4 * i < W) *reinterpret_cast<float4*>(o + 4 * i) = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
    } else {
#pragma unroll
      for (int p = 0; p < PX; ++p)
        if (xs + p < W) o[p] = r[p];
    }
  }
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
// mode 0: dout = d f1, fo = f2 ; mode 1: dout = d f2, fo = f1.  blockIdx.z = b * nmodes + slot.
template <class T, int CR, bool VEC>
__global__ void __launch_bounds__(T::THREADS, 2)
corr_bwd_tiled(const float* __restrict__ g, const float* __restrict__ oact, const float* __restrict__ f1,
               const float* __restrict__ f2, float* __restrict__ df1, float* __restrict__ df2, int C, int H, int W,
               long long g_bstride, float inv_c, float slope, int nmodes, int first_mode) {
  constexpr int D = T::D, PX = T::PX, ND = T::ND, TH = T::TH, TW = T::TW, CC = T::CC;
  constexpr int S1 = T::S1, S2 = T::S2, F2H = T::F2H, F2W = T::F2W, WIN = T::WIN;
  static_assert(CC % CR == 0, "CC must be a multiple of CR");
  extern __shared__ __align__(16) float smem[];
  float* fos = smem;                   // [CC][F2H][S2]
  float* red = smem + CC * F2H * S2;   // [ND][CR][TH][S1]

  const int tid = threadIdx.x;
  const int lane = tid % T::LANES;
  const int tx = lane % T::TXT, ty = lane / T::TXT;
  const int dyi = tid / T::LANES;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int b = blockIdx.z / nmodes;
  const int mode = first_mode + (blockIdx.z - b * nmodes);
  const float* fo = (mode == 0 ? f2 : f1) + (size_t)b * C * H * W;
  float* dout = (mode == 0 ? df1 : df2) + (size_t)b * C * H * W;
  const size_t gb = (g_bstride ? (size_t)g_bstride : (size_t)ND * ND * H * W) * b;

  // 81 per-pixel coefficients of this thread's dy row, kept in registers for the whole kernel
  float G[ND][PX];
  {
    const int y = y0 + ty, xs = x0 + tx * PX;
#pragma unroll
    for (int dx = 0; dx < ND; ++dx) {
      // mode 0: plane k(dy,dx) at (y, x) ; mode 1: plane k(-dy,-dx) at (y+dy, x+dx)
      const int k = mode == 0 ? dyi * ND + dx : (2 * D - dyi) * ND + (2 * D - dx);
      const int sy = mode == 0 ? y : y + dyi - D;
      const int sx0 = mode == 0 ? xs : xs + dx - D;
      const size_t off = gb + ((size_t)k * H + sy) * W;
#pragma unroll
      for (int p = 0; p < PX; ++p) {
        const int sx = sx0 + p;
        float v = 0.f;
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
          v = __ldg(g + off + sx);
          if (oact != nullptr && !(__ldg(oact + off + sx) > 0.f)) v *= slope;
        }
        G[dx][p] = v;
      }
    }
  }

  for (int c0 = 0; c0 < C; c0 += CC) {
    __syncthreads();
    stage_box<CC, F2H, F2W, S2, T::THREADS, VEC>(fos, fo, c0, C, H, W, y0 - D, x0 - D, 0.f, 1.f, false);
    __syncthreads();
    const float* pw = fos + (ty + dyi) * S2 + tx * PX;
#pragma unroll 1
    for (int r0 = 0; r0 < CC; r0 += CR) {
      if (c0 + r0 >= C) break;  // uniform across the block
#pragma unroll
      for (int cr = 0; cr < CR; ++cr) {
        float w[WIN], part[PX];
#pragma unroll
        for (int i = 0; i < WIN / 4; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(pw + (r0 + cr) * F2H * S2 + 4 * i);
          w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int p = 0; p < PX; ++p) part[p] = 0.f;
#pragma unroll
        for (int dx = 0; dx < ND; ++dx)
#pragma unroll
          for (int p = 0; p < PX; ++p) part[p] = fmaf(G[dx][p], w[p + dx], part[p]);
        float* rp = red + ((dyi * CR + cr) * TH + ty) * S1 + tx * PX;
#pragma unroll
        for (int i = 0; i < PX / 4; ++i)
          *reinterpret_cast<float4*>(rp + 4 * i) = make_float4(part[4 * i], part[4 * i + 1], part[4 * i + 2], part[4 * i + 3]);
      }
      __syncthreads();
      // cross-dy reduction: CR*TH*TW/4 float4 outputs
      constexpr int OUT4 = CR * TH * TW / 4;
      for (int i = tid; i < OUT4; i += T::THREADS) {
        const int x4 = i % (TW / 4), row = i / (TW / 4);  // row = cr*TH + ry
        const int cr = row / TH, ry = row - cr * TH;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          const float4 v = *reinterpret_cast<const float4*>(red + ((k * CR + cr) * TH + ry) * S1 + x4 * 4);
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        const int c = c0 + r0 + cr, y = y0 + ry, x = x0 + x4 * 4;
        if (c < C && y < H && x < W) {
          float* o = dout + ((size_t)c * H + y) * W + x;
          if (VEC) {
            *reinterpret_cast<float4*>(o) = make_float4(s.x * inv_c, s.y * inv_c, s.z * inv_c, s.w * inv_c);
          } else {
            o[0] = s.x * inv_c;
            if (x + 1 < W) o[1] = s.y * inv_c;
            if (x + 2 < W) o[2] = s.z * inv_c;
            if (x + 3 < W) o[3] = s.w * inv_c;
          }
        }
      }
      __syncthreads();
    }
  }
}

// generic-displacement backward: one thread per (b, c, y, x) element of d f1 / d f2.
__global__ void __launch_bounds__(128)
corr_bwd_generic(const float* __restrict__ g, const float* __restrict__ oact, const float* __restrict__ f1,
                 const float* __restrict__ f2, float* __restrict__ df1, float* __restrict__ df2, int C, int H, int W, int d,
                 long long g_bstride, float inv_c, float slope) {
  const int nd = 2 * d + 1;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= H * W) return;
  const int y = pix / W, x = pix - y * W;
  const int c = blockIdx.y, b = blockIdx.z;
  const size_t gb = (g_bstride ? (size_t)g_bstride : (size_t)nd * nd * H * W) * b;
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
        if (oact != nullptr && !(__ldg(oact + go) > 0.f)) gv *= slope;
        a1 = fmaf(gv, __ldg(f2 + fb + (size_t)yy * W + xx), a1);
      }
      // d f2: g[k, y-dy, x-dx] * f1[y-dy, x-dx]
      const int ys = y - dy, xs = x - dx;
      if (df2 != nullptr && ys >= 0 && ys < H && xs >= 0 && xs < W) {
        const size_t go = gb + ((size_t)k * H + ys) * W + xs;
        float gv = __ldg(g + go);
        if (oact != nullptr && !(__ldg(oact + go) > 0.f)) gv *= slope;
        a2 = fmaf(gv, __ldg(f1 + fb + (size_t)ys * W + xs), a2);
      }
    }
  }
  if (df1 != nullptr) df1[fb + pix] = a1 * inv_c;
  if (df2 != nullptr) df2[fb + pix] = a2 * inv_c;
}

using Tile4 = CorrTile<4, 8, 4, 8, 16>;
constexpr int BWD_CR = 4;

template <class K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

}  // namespace

extern "C" int ocf_corr_fwd(const float* f1, const float* f2, float* out, int B, int C, int H, int W, int d,
                            long long out_bstride, float leaky_slope, const float* norm, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(f1); OCF_REQUIRE_PTR(f2); OCF_REQUIRE_PTR(out);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(d >= 0 && d <= OCF_MAX_DISPLACEMENT, OCF_EUNSUPPORTED);
  const long long nd = 2 * d + 1;
  OCF_REQUIRE(out_bstride == 0 || out_bstride >= nd * nd * H * W, OCF_ESHAPE);
  OCF_REQUIRE(B <= 65535, OCF_EUNSUPPORTED);
  cudaStream_t s = ocf_cast_stream(stream);
  const float inv_c = 1.0f / (float)C;
  if (d == 4) {
    using T = Tile4;
    const size_t smem = sizeof(float) * T::CC * (T::TH * T::S1 + T::F2H * T::S2);
    dim3 grid((W + T::TW - 1) / T::TW, (H + T::TH - 1) / T::TH, B);
    const bool vec = (W % 4 == 0) && ocf_aligned16(f1) && ocf_aligned16(f2) && ocf_aligned16(out) && (out_bstride % 4 == 0);
    if (vec) {
      if (int e = set_smem(corr_fwd_tiled<T, true>, smem)) return e;
      corr_fwd_tiled<T, true><<<grid, T::THREADS, smem, s>>>(f1, f2, out, C, H, W, out_bstride, inv_c, leaky_slope, norm);
    } else {
      if (int e = set_smem(corr_fwd_tiled<T, false>, smem)) return e;
      corr_fwd_tiled<T, false><<<grid, T::THREADS, smem, s>>>(f1, f2, out, C, H, W, out_bstride, inv_c, leaky_slope, norm);
    }
  } else {
    dim3 grid((H * W + 127) / 128, (unsigned)nd, B);
    corr_fwd_generic<<<grid, 128, 0, s>>>(f1, f2, out, C, H, W, d, out_bstride, inv_c, leaky_slope, norm);
  }
  return ocf_launch_status();
}

extern "C" int ocf_corr_bwd(const float* grad_out, const float* out_act, const float* f1, const float* f2, float* df1,
                            float* df2, int B, int C, int H, int W, int d, long long g_bstride, float leaky_slope,
                            ocf_stream_t stream) {
  OCF_REQUIRE_PTR(grad_out); OCF_REQUIRE_PTR(f1); OCF_REQUIRE_PTR(f2);
  OCF_REQUIRE(df1 != nullptr || df2 != nullptr, OCF_ENULL);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(d >= 0 && d <= OCF_MAX_DISPLACEMENT, OCF_EUNSUPPORTED);
  const long long nd = 2 * d + 1;
  OCF_REQUIRE(g_bstride == 0 || g_bstride >= nd * nd * H * W, OCF_ESHAPE);
  OCF_REQUIRE(B <= 32767 && C <= 65535, OCF_EUNSUPPORTED);
  cudaStream_t s = ocf_cast_stream(stream);
  const float inv_c = 1.0f / (float)C;
  if (d == 4) {
    using T = Tile4;
    const size_t smem = sizeof(float) * (T::CC * T::F2H * T::S2 + T::ND * BWD_CR * T::TH * T::S1);
    const int nmodes = (df1 != nullptr && df2 != nullptr) ? 2 : 1;
    const int first = df1 != nullptr ? 0 : 1;
    dim3 grid((W + T::TW - 1) / T::TW, (H + T::TH - 1) / T::TH, B * nmodes);
    const bool vec = (W % 4 == 0) && ocf_aligned16(f1) && ocf_aligned16(f2) && (df1 == nullptr || ocf_aligned16(df1)) &&
                     (df2 == nullptr || ocf_aligned16(df2));
    if (vec) {
      if (int e = set_smem(corr_bwd_tiled<T, BWD_CR, true>, smem)) return e;
      corr_bwd_tiled<T, BWD_CR, true><<<grid, T::THREADS, smem, s>>>(grad_out, out_act, f1, f2, df1, df2, C, H, W, g_bstride,
                                                                    inv_c, leaky_slope, nmodes, first);
    } else {
      if (int e = set_smem(corr_bwd_tiled<T, BWD_CR, false>, smem)) return e;
      corr_bwd_tiled<T, BWD_CR, false><<<grid, T::THREADS, smem, s>>>(grad_out, out_act, f1, f2, df1, df2, C, H, W, g_bstride,
                                                                     inv_c, leaky_slope, nmodes, first);
    }
  } else {
    dim3 grid((H * W + 127) / 128, C, B);
    corr_bwd_generic<<<grid, 128, 0, s>>>(grad_out, out_act, f1, f2, df1, df2, C, H, W, d, g_bstride, inv_c, leaky_slope);
  }
  return ocf_launch_status();
}
