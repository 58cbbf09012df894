// SSIM between two image batches, forward and backward, fp32 NCHW.
//
// Replaces _ssim / ssim / SSIM of the reference (inpainting_metrics/ssim/ssim.py:7-37, :39-75): five depthwise
// F.conv2d calls with a ws x ws Gaussian (sigma 1.5, zero padding ws//2) + ~15 elementwise launches => one launch
// forward, one backward.  C1 = 0.01^2, C2 = 0.03^2.  Even windows (the reference uses ws = 4 at
// inpainting_metrics/__init__.py:23) produce a (H+1) x (W+1) map, exactly like conv2d with padding ws//2.
//
// The Gaussian is separable (the reference builds its 2-D window as the outer product of the normalised 1-D one), so a
// tile is blurred with a horizontal and a vertical pass through shared memory: 2*ws taps per statistic instead of ws^2.
// Forward statistics per map pixel: mu1, mu2, E[x^2], E[y^2], E[xy] -> S; optionally the four partial derivatives the
// backward needs (dS/d blur(x), dS/d blur(y), dS/d blur(x^2) == dS/d blur(y^2), dS/d blur(xy)).
// Backward: d img1 = adj(Kx) + 2*img1*adj(Kq) + img2*adj(Kxy) (and symmetrically for img2), adj = the transposed blur.
#include "common.cuh"

#include <math.h>

namespace {

constexpr int SSIM_MAXW = 15;
constexpr int TW = 32, TH = 16, NTHR = 256;
constexpr int HALO = SSIM_MAXW - 1;
constexpr int SW = TW + HALO + 2;  // shared row stride (48: rows stay 16-byte aligned, column reads are conflict-free)

struct Taps {
  float g[SSIM_MAXW + 1];
};

__device__ __forceinline__ float ssim_value(float mu1, float mu2, float exx, float eyy, float exy, float* kx, float* ky, float* kq,
                                            float* kxy) {
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
  const float s1 = exx - mu1_sq, s2 = eyy - mu2_sq, s12 = exy - mu12;
  const float N1 = 2.f * mu12 + C1, N2 = 2.f * s12 + C2;
  const float D1 = mu1_sq + mu2_sq + C1, D2 = s1 + s2 + C2;
  const float inv = 1.0f / (D1 * D2);
  const float S = (N1 * N2) * inv;
  if (kx != nullptr) {
    const float a = N2 * inv, b = N1 * inv, c = -S / D1, d = -S / D2;
    // blur(x) enters N1 (2*mu2), N2 through sigma12 (-mu2 * 2), D1 (2*mu1), D2 through sigma1^2 (-2*mu1)
    *kx = 2.f * mu2 * (a - b) + 2.f * mu1 * (c - d);
    *ky = 2.f * mu1 * (a - b) + 2.f * mu2 * (c - d);
    *kq = d;         // d S / d blur(x^2) == d S / d blur(y^2)
    *kxy = 2.f * b;  // d S / d blur(xy)
  }
  return S;
}

// grid (ceil(Wo/TW), ceil(Ho/TH), B*C); coef layout [B*C][4][Ho][Wo]
__global__ void __launch_bounds__(NTHR)
ssim_fwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2, double* __restrict__ sums, float* __restrict__ coef,
                int C, int H, int W, int Ho, int Wo, int ws, Taps taps) {
  __shared__ float xin[TH + HALO][SW], yin[TH + HALO][SW];
  __shared__ float hs[5][TH + HALO][TW];
  const int p = ws / 2;
  const int bc = blockIdx.z;
  const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
  const float* x = img1 + (size_t)bc * H * W;
  const float* y = img2 + (size_t)bc * H * W;
  const int rows = TH + ws - 1, cols = TW + ws - 1;
  // input window of the tile: rows oy0 - p .. oy0 - p + rows - 1
  for (int i = threadIdx.x; i < rows * cols; i += NTHR) {
    const int r = i / cols, c = i - r * cols;
    const int gy = oy0 - p + r, gx = ox0 - p + c;
    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
    xin[r][c] = ok ? __ldg(x + (size_t)gy * W + gx) : 0.f;
    yin[r][c] = ok ? __ldg(y + (size_t)gy * W + gx) : 0.f;
  }
  __syncthreads();
  // horizontal pass: 5 statistics for every staged row
  for (int i = threadIdx.x; i < rows * TW; i += NTHR) {
    const int r = i / TW, c = i - r * TW;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
    for (int k = 0; k < ws; ++k) {
      const float g = taps.g[k], u = xin[r][c + k], v = yin[r][c + k];
      a0 = fmaf(g, u, a0); a1 = fmaf(g, v, a1);
      a2 = fmaf(g, u * u, a2); a3 = fmaf(g, v * v, a3); a4 = fmaf(g, u * v, a4);
    }
    hs[0][r][c] = a0; hs[1][r][c] = a1; hs[2][r][c] = a2; hs[3][r][c] = a3; hs[4][r][c] = a4;
  }
  __syncthreads();
  // vertical pass + SSIM
  float acc[1] = {0.f};
  const size_t plane = (size_t)Ho * Wo;
  for (int i = threadIdx.x; i < TH * TW; i += NTHR) {
    const int r = i / TW, c = i - r * TW;
    const int oy = oy0 + r, ox = ox0 + c;
    if (oy >= Ho || ox >= Wo) continue;
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < ws; ++k) {
      const float g = taps.g[k];
#pragma unroll
      for (int q = 0; q < 5; ++q) m[q] = fmaf(g, hs[q][r + k][c], m[q]);
    }
    float kx, ky, kq, kxy;
    const bool want = coef != nullptr;
    const float S = ssim_value(m[0], m[1], m[2], m[3], m[4], want ? &kx : nullptr, &ky, &kq, &kxy);
    acc[0] += S;
    if (want) {
      float* o = coef + (size_t)bc * 4 * plane + (size_t)oy * Wo + ox;
      o[0] = kx; o[plane] = ky; o[2 * plane] = kq; o[3 * plane] = kxy;
    }
  }
  ocf_block_accumulate<1>(acc, sums + bc / C);
}

// grid (ceil(W/TW), ceil(H/TH), B*C): d img1 / d img2 tile; scale[b] multiplies everything (upstream gradient / numel)
__global__ void __launch_bounds__(NTHR)
ssim_bwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2, const float* __restrict__ coef,
                const float* __restrict__ scale, float* __restrict__ d1, float* __restrict__ d2, int C, int H, int W, int Ho, int Wo,
                int ws, Taps taps) {
  __shared__ float kin[4][TH + HALO][SW];
  __shared__ float hs[4][TH + HALO][TW];
  const int p = ws / 2;
  const int bc = blockIdx.z;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const size_t plane = (size_t)Ho * Wo;
  const float* kb = coef + (size_t)bc * 4 * plane;
  const int rows = TH + ws - 1, cols = TW + ws - 1;
  // image pixel (y, x) receives map pixel (oy, ox) with weight g[y - oy + p] * g[x - ox + p], oy in [y + p - ws + 1, y + p]
  const int oyb = y0 + p - ws + 1, oxb = x0 + p - ws + 1;
  for (int i = threadIdx.x; i < rows * cols; i += NTHR) {
    const int r = i / cols, c = i - r * cols;
    const int oy = oyb + r, ox = oxb + c;
    const bool ok = oy >= 0 && oy < Ho && ox >= 0 && ox < Wo;
    const size_t off = (size_t)oy * Wo + ox;
#pragma unroll
    for (int q = 0; q < 4; ++q) kin[q][r][c] = ok ? __ldg(kb + q * plane + off) : 0.f;
  }
  __syncthreads();
  // horizontal adjoint: T[r][c] = sum_k K[r][c + k] * g[ws - 1 - k]   (x = x0 + c, ox = oxb + c + k => tap x - ox + p = ws-1-k)
  for (int i = threadIdx.x; i < rows * TW; i += NTHR) {
    const int r = i / TW, c = i - r * TW;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < ws; ++k) {
      const float g = taps.g[ws - 1 - k];
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] = fmaf(g, kin[q][r][c + k], a[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) hs[q][r][c] = a[q];
  }
  __syncthreads();
  const float sc = __ldg(scale + bc / C);
  const float* x = img1 + (size_t)bc * H * W;
  const float* y = img2 + (size_t)bc * H * W;
  for (int i = threadIdx.x; i < TH * TW; i += NTHR) {
    const int r = i / TW, c = i - r * TW;
    const int gy = y0 + r, gx = x0 + c;
    if (gy >= H || gx >= W) continue;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < ws; ++k) {
      const float g = taps.g[ws - 1 - k];
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] = fmaf(g, hs[q][r + k][c], a[q]);
    }
    const size_t off = (size_t)gy * W + gx;
    const float u = __ldg(x + off), v = __ldg(y + off);
    if (d1 != nullptr) d1[(size_t)bc * H * W + off] = sc * (a[0] + 2.f * u * a[2] + v * a[3]);
    if (d2 != nullptr) d2[(size_t)bc * H * W + off] = sc * (a[1] + 2.f * v * a[2] + u * a[3]);
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
  dim3 grid((Wo + TW - 1) / TW, (Ho + TH - 1) / TH, B * C);
  ssim_fwd_kernel<<<grid, NTHR, 0, s>>>(img1, img2, sums, coef, C, H, W, Ho, Wo, window, make_taps(window));
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
  dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B * C);
  ssim_bwd_kernel<<<grid, NTHR, 0, ocf_cast_stream(stream)>>>(img1, img2, coef, scale, d_img1, d_img2, C, H, W, Ho, Wo, window,
                                                              make_taps(window));
  return ocf_launch_status();
}
