// SSIM between two image batches, forward and backward, fp32 NCHW.
//
// Replaces _ssim / ssim / SSIM of the reference (inpainting_metrics/ssim/ssim.py:7-37, :39-75): five depthwise
// F.conv2d calls with a ws x ws Gaussian (sigma 1.5, zero padding ws//2) + ~15 elementwise launches => one launch
// forward, one backward.  C1 = 0.01^2, C2 = 0.03^2.  Even windows (the reference uses ws = 4 at
// inpainting_metrics/__init__.py:23) produce a (H+1) x (W+1) map, exactly like conv2d with padding ws//2.
//
// The Gaussian is separable (the reference builds its 2-D window as the outer product of the normalised 1-D one), so a
// tile is blurred with a horizontal and a vertical pass through shared memory: 2*ws taps per statistic instead of ws^2.
// The window size is a template parameter (1..15 instantiated): both passes are fully unrolled register sliding windows
// (4 outputs per thread share ws+3 shared-memory reads) with the taps as constant-bank operands.
// Forward statistics per map pixel: mu1, mu2, E[x^2], E[y^2], E[xy] -> S; optionally the four partial derivatives the
// backward needs (dS/d blur(x), dS/d blur(y), dS/d blur(x^2) == dS/d blur(y^2), dS/d blur(xy)).
// Backward: d img1 = adj(Kx) + 2*img1*adj(Kq) + img2*adj(Kxy) (and symmetrically for img2), adj = the transposed blur.
#include "common.cuh"

#include <math.h>

namespace {

constexpr int SSIM_MAXW = 15;
constexpr int TW = 32, NTHR = 256;
constexpr int PXT = 4;                  // horizontal pass: consecutive output columns per thread
constexpr int SW = TW + SSIM_MAXW + 1;  // shared row stride (48 floats: rows stay 16-byte aligned)

struct Taps {
  float g[SSIM_MAXW + 1];
};

__device__ __forceinline__ float ssim_value(float mu1, float mu2, float exx, float eyy, float exy, bool want, float* kx, float* ky,
                                            float* kq, float* kxy) {
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
  const float s1 = exx - mu1_sq, s2 = eyy - mu2_sq, s12 = exy - mu12;
  const float N1 = 2.f * mu12 + C1, N2 = 2.f * s12 + C2;
  const float D1 = mu1_sq + mu2_sq + C1, D2 = s1 + s2 + C2;
  const float inv = 1.0f / (D1 * D2);
  const float S = (N1 * N2) * inv;
  if (want) {
    const float a = N2 * inv, b = N1 * inv, c = -S / D1, d = -S / D2;
    // blur(x) enters N1 (2*mu2), N2 through sigma12 (-mu2 * 2), D1 (2*mu1), D2 through sigma1^2 (-2*mu1)
    *kx = 2.f * mu2 * (a - b) + 2.f * mu1 * (c - d);
    *ky = 2.f * mu1 * (a - b) + 2.f * mu2 * (c - d);
    *kq = d;         // d S / d blur(x^2) == d S / d blur(y^2)
    *kxy = 2.f * b;  // d S / d blur(xy)
  }
  return S;
}

// Row-wise separable blur of NQ staged planes: src[q][r][0 .. TW+WS-2] -> dst[q][r][0 .. TW-1] for r < rows.
// A thread owns PXT consecutive output columns of one row and slides over a register window of PXT+WS-1 inputs
// (WS is a compile-time constant: fully unrolled, the taps are constant-bank operands of the FFMAs).
// REV: use the taps in reverse order (the adjoint blur of the backward pass).  XFORM: planes 2..4 of the forward pass are
// the products x^2, y^2, xy formed on the fly from the two staged images.
template <int WS, int NQ, bool REV, bool XFORM, int ROWS>
__device__ __forceinline__ void blur_rows(const float (*src)[ROWS][SW], float (*dst)[ROWS][TW], int rows, const Taps& taps) {
  constexpr int WIN = PXT + WS - 1;
  constexpr int NSRC = XFORM ? 2 : NQ;
  for (int i = threadIdx.x; i < rows * (TW / PXT); i += NTHR) {
    const int r = i / (TW / PXT), c0 = (i - r * (TW / PXT)) * PXT;
    float w[NSRC][WIN];
#pragma unroll
    for (int q = 0; q < NSRC; ++q)
#pragma unroll
      for (int j = 0; j < WIN; ++j) w[q][j] = src[q][r][c0 + j];
    float acc[NQ][PXT];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int p = 0; p < PXT; ++p) acc[q][p] = 0.f;
#pragma unroll
    for (int k = 0; k < WS; ++k) {
      const float g = taps.g[REV ? WS - 1 - k : k];
#pragma unroll
      for (int p = 0; p < PXT; ++p) {
        if (XFORM) {
          const float u = w[0][p + k], v = w[1][p + k];
          acc[0][p] = fmaf(g, u, acc[0][p]); acc[1][p] = fmaf(g, v, acc[1][p]);
          acc[2][p] = fmaf(g, u * u, acc[2][p]); acc[3][p] = fmaf(g, v * v, acc[3][p]); acc[4][p] = fmaf(g, u * v, acc[4][p]);
        } else {
#pragma unroll
          for (int q = 0; q < NQ; ++q) acc[q][p] = fmaf(g, w[q][p + k], acc[q][p]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) *reinterpret_cast<float4*>(&dst[q][r][c0]) = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
  }
}

// Column-wise blur: thread (column c, rows r0 .. r0+RPT-1) slides over RPT+WS-1 rows of dst-type planes; result in out[q][j].
template <int WS, int NQ, bool REV, int RPT, int ROWS>
__device__ __forceinline__ void blur_cols(const float (*hs)[ROWS][TW], int r0, int c, const Taps& taps, float (&out)[NQ][RPT]) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    float w[RPT + WS - 1];
#pragma unroll
    for (int j = 0; j < RPT + WS - 1; ++j) w[j] = hs[q][r0 + j][c];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < WS; ++k) a = fmaf(taps.g[REV ? WS - 1 - k : k], w[j + k], a);
      out[q][j] = a;
    }
  }
}

// grid (ceil(Wo/TW), ceil(Ho/TH), B*C); coef layout [B*C][4][Ho][Wo]
template <int WS, int TH>
__global__ void __launch_bounds__(NTHR)
ssim_fwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2, double* __restrict__ sums, float* __restrict__ coef,
                int C, int H, int W, int Ho, int Wo, const __grid_constant__ Taps taps) {
  constexpr int ROWS = TH + WS - 1, COLS = TW + WS - 1, RPT = TH * TW / NTHR;
  __shared__ __align__(16) float in[2][ROWS][SW];
  __shared__ __align__(16) float hs[5][ROWS][TW];
  constexpr int p = WS / 2;
  const int bc = blockIdx.z;
  const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
  const float* x = img1 + (size_t)bc * H * W;
  const float* y = img2 + (size_t)bc * H * W;
  // input window of the tile: rows oy0 - p .. oy0 - p + ROWS - 1, zero outside the image (F.conv2d padding)
  for (int i = threadIdx.x; i < ROWS * COLS; i += NTHR) {
    const int r = i / COLS, c = i - r * COLS;
    const int gy = oy0 - p + r, gx = ox0 - p + c;
    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
    in[0][r][c] = ok ? __ldg(x + (size_t)gy * W + gx) : 0.f;
    in[1][r][c] = ok ? __ldg(y + (size_t)gy * W + gx) : 0.f;
  }
  __syncthreads();
  blur_rows<WS, 5, false, true, ROWS>(in, hs, ROWS, taps);
  __syncthreads();
  float acc[1] = {0.f};
  const size_t plane = (size_t)Ho * Wo;
  const int c = threadIdx.x % TW, r0 = (threadIdx.x / TW) * RPT;
  float m[5][RPT];
  blur_cols<WS, 5, false, RPT, ROWS>(hs, r0, c, taps, m);
  const bool want = coef != nullptr;
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int oy = oy0 + r0 + j, ox = ox0 + c;
    if (oy >= Ho || ox >= Wo) continue;
    float kx, ky, kq, kxy;
    acc[0] += ssim_value(m[0][j], m[1][j], m[2][j], m[3][j], m[4][j], want, &kx, &ky, &kq, &kxy);
    if (want) {
      float* o = coef + (size_t)bc * 4 * plane + (size_t)oy * Wo + ox;
      o[0] = kx; o[plane] = ky; o[2 * plane] = kq; o[3 * plane] = kxy;
    }
  }
  ocf_block_accumulate<1>(acc, sums + bc / C);
}

// grid (ceil(W/TW), ceil(H/TH), B*C): d img1 / d img2 tile; scale[b] multiplies everything (upstream gradient / numel)
template <int WS, int TH>
__global__ void __launch_bounds__(NTHR)
ssim_bwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2, const float* __restrict__ coef,
                const float* __restrict__ scale, float* __restrict__ d1, float* __restrict__ d2, int C, int H, int W, int Ho, int Wo,
                const __grid_constant__ Taps taps) {
  constexpr int ROWS = TH + WS - 1, COLS = TW + WS - 1, RPT = TH * TW / NTHR;
  __shared__ __align__(16) float kin[4][ROWS][SW];
  __shared__ __align__(16) float hs[4][ROWS][TW];
  constexpr int p = WS / 2;
  const int bc = blockIdx.z;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const size_t plane = (size_t)Ho * Wo;
  const float* kb = coef + (size_t)bc * 4 * plane;
  // image pixel (y, x) receives map pixel (oy, ox) with weight g[y - oy + p] * g[x - ox + p], oy in [y + p - WS + 1, y + p]
  const int oyb = y0 + p - WS + 1, oxb = x0 + p - WS + 1;
  for (int i = threadIdx.x; i < ROWS * COLS; i += NTHR) {
    const int r = i / COLS, c = i - r * COLS;
    const int oy = oyb + r, ox = oxb + c;
    const bool ok = oy >= 0 && oy < Ho && ox >= 0 && ox < Wo;
    const size_t off = (size_t)oy * Wo + ox;
#pragma unroll
    for (int q = 0; q < 4; ++q) kin[q][r][c] = ok ? __ldg(kb + q * plane + off) : 0.f;
  }
  __syncthreads();
  // horizontal adjoint: T[r][c] = sum_k K[r][c + k] * g[WS - 1 - k]   (x = x0 + c, ox = oxb + c + k => tap x - ox + p = WS-1-k)
  blur_rows<WS, 4, true, false, ROWS>(kin, hs, ROWS, taps);
  __syncthreads();
  const float sc = __ldg(scale + bc / C);
  const float* x = img1 + (size_t)bc * H * W;
  const float* y = img2 + (size_t)bc * H * W;
  const int c = threadIdx.x % TW, r0 = (threadIdx.x / TW) * RPT;
  float a[4][RPT];
  blur_cols<WS, 4, true, RPT, ROWS>(hs, r0, c, taps, a);
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int gy = y0 + r0 + j, gx = x0 + c;
    if (gy >= H || gx >= W) continue;
    const size_t off = (size_t)gy * W + gx;
    const float u = __ldg(x + off), v = __ldg(y + off);
    if (d1 != nullptr) d1[(size_t)bc * H * W + off] = sc * (a[0][j] + 2.f * u * a[2][j] + v * a[3][j]);
    if (d2 != nullptr) d2[(size_t)bc * H * W + off] = sc * (a[1][j] + 2.f * v * a[2][j] + u * a[3][j]);
  }
}

// gaussian(window_size, 1.5) of the reference: fp32 taps exp(-(i - ws//2)^2 / (2 sigma^2)), normalised by their fp32 sum
Taps make_taps(int ws) {
  Taps t;
  float s = 0.f;
  for (int i = 0; i < ws; ++i) {
    const double d = (double)(i - ws / 2);
    t.g[i] = (float)exp(-(d * d) / (2.0 * 1.5 * 1.5));
    s += t.g[i];
  }
  for (int i = 0; i < ws; ++i) t.g[i] = t.g[i] / s;
  for (int i = ws; i <= SSIM_MAXW; ++i) t.g[i] = 0.f;
  return t;
}

constexpr int FWD_TH = 32, BWD_TH = 16;  // tile heights: 42 KB / 33 KB of static shared memory at WS = 11

template <int WS>
void launch_fwd(dim3 grid, cudaStream_t s, const float* a, const float* b, double* sums, float* coef, int C, int H, int W, int Ho, int Wo) {
  ssim_fwd_kernel<WS, FWD_TH><<<grid, NTHR, 0, s>>>(a, b, sums, coef, C, H, W, Ho, Wo, make_taps(WS));
}
template <int WS>
void launch_bwd(dim3 grid, cudaStream_t s, const float* a, const float* b, const float* coef, const float* scale, float* d1, float* d2,
                int C, int H, int W, int Ho, int Wo) {
  ssim_bwd_kernel<WS, BWD_TH><<<grid, NTHR, 0, s>>>(a, b, coef, scale, d1, d2, C, H, W, Ho, Wo, make_taps(WS));
}

#define OCF_SSIM_DISPATCH(ws, CALL) \
  switch (ws) {                     \
    case 1: CALL(1); break;   case 2: CALL(2); break;   case 3: CALL(3); break;   case 4: CALL(4); break;   case 5: CALL(5); break;    \
    case 6: CALL(6); break;   case 7: CALL(7); break;   case 8: CALL(8); break;   case 9: CALL(9); break;   case 10: CALL(10); break;  \
    case 11: CALL(11); break; case 12: CALL(12); break; case 13: CALL(13); break; case 14: CALL(14); break; default: CALL(15); break; \
  }

}  // namespace

extern "C" int ocf_ssim_fwd(const float* img1, const float* img2, double* sums, float* coef, int B, int C, int H, int W, int window,
                            ocf_stream_t stream) {
  OCF_REQUIRE_PTR(img1); OCF_REQUIRE_PTR(img2); OCF_REQUIRE_PTR(sums);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(window >= 1 && window <= SSIM_MAXW, OCF_EUNSUPPORTED);
  OCF_REQUIRE((long long)B * C <= 65535, OCF_EUNSUPPORTED);
  cudaStream_t s = ocf_cast_stream(stream);
  const int Ho = H + 2 * (window / 2) - window + 1, Wo = W + 2 * (window / 2) - window + 1;
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * B, s);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((Wo + TW - 1) / TW, (Ho + FWD_TH - 1) / FWD_TH, B * C);
#define OCF_CALL(WS) launch_fwd<WS>(grid, s, img1, img2, sums, coef, C, H, W, Ho, Wo)
  OCF_SSIM_DISPATCH(window, OCF_CALL)
#undef OCF_CALL
  return ocf_launch_status();
}

extern "C" int ocf_ssim_bwd(const float* img1, const float* img2, const float* coef, const float* scale, float* d_img1, float* d_img2,
                            int B, int C, int H, int W, int window, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(img1); OCF_REQUIRE_PTR(img2); OCF_REQUIRE_PTR(coef); OCF_REQUIRE_PTR(scale);
  OCF_REQUIRE(d_img1 != nullptr || d_img2 != nullptr, OCF_ENULL);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(window >= 1 && window <= SSIM_MAXW, OCF_EUNSUPPORTED);
  OCF_REQUIRE((long long)B * C <= 65535, OCF_EUNSUPPORTED);
  const int Ho = H + 2 * (window / 2) - window + 1, Wo = W + 2 * (window / 2) - window + 1;
  dim3 grid((W + TW - 1) / TW, (H + BWD_TH - 1) / BWD_TH, B * C);
  cudaStream_t s = ocf_cast_stream(stream);
#define OCF_CALL(WS) launch_bwd<WS>(grid, s, img1, img2, coef, scale, d_img1, d_img2, C, H, W, Ho, Wo)
  OCF_SSIM_DISPATCH(window, OCF_CALL)
#undef OCF_CALL
  return ocf_launch_status();
}
