// Soft census (ternary) photometric term, forward and backward, fp32, NCHW.
//
// BASELINE.json's north_star lists a census term among the photometric losses; the reference defines none
// (SURVEY.md section 8a-14), so this follows the published UnFlow soft census (grey*255, (2m+1)^2 patch,
// t = u / sqrt(0.81 + u^2), soft Hamming dt^2 / (0.1 + dt^2), m-pixel border masked) with the occlusion weighting of
// photometric_error (models/model.py:37-46).  PARITY UNPINNED: checked against oracle/ocflow_oracle.py::census_loss only.
//
// Only interior pixels (m <= y < H-m, m <= x < W-m) carry weight, and their patches lie inside the image, so padding
// never enters.  One CTA = a 16 x 32 pixel tile; the grey images of the tile + halo (m forward, 2m backward) are built
// once in shared memory from coalesced loads of the C channels; every patch access after that is a shared-memory read.
#include "common.cuh"

namespace {

constexpr int CT_W = 32, CT_H = 16, CT_THREADS = 256;

__device__ __forceinline__ float grey_weight(int C, int c) {
  if (C == 3) return c == 0 ? 0.2989f : (c == 1 ? 0.5870f : 0.1140f);
  return 1.0f / (float)C;
}

// grey*255 of pixel (y, x) of batch item b; 0 outside the image (never used by a weighted pixel)
__device__ __forceinline__ float grey255(const float* __restrict__ img, int C, int H, int W, int y, int x) {
  if (y < 0 || y >= H || x < 0 || x >= W) return 0.f;
  const size_t HW = (size_t)H * W;
  const float* p = img + (size_t)y * W + x;
  float g;
  if (C == 3) {
    // same association as the oracle: (r*0.2989 + g*0.5870) + b*0.1140, products rounded separately
    g = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(p), 0.2989f), __fmul_rn(__ldg(p + HW), 0.5870f)), __fmul_rn(__ldg(p + 2 * HW), 0.1140f));
  } else {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += __ldg(p + c * HW);
    g = s / (float)C;
  }
  return g * 255.0f;
}

// the product is rounded on its own (no FMA contraction with the subtraction that follows), so identical inputs give exactly 0
__device__ __forceinline__ float soft_sign(float u) { return __fmul_rn(u, rsqrtf(fmaf(u, u, 0.81f))); }

template <int M, int HALO>
__device__ __forceinline__ void stage_grey(float (*g1)[CT_W + 2 * HALO], float (*g2)[CT_W + 2 * HALO], const float* __restrict__ a,
                                           const float* __restrict__ b, int C, int H, int W, int y0, int x0) {
  constexpr int SW = CT_W + 2 * HALO, SH = CT_H + 2 * HALO;
  for (int i = threadIdx.x; i < SW * SH; i += CT_THREADS) {
    const int r = i / SW, c = i - r * SW;
    g1[r][c] = grey255(a, C, H, W, y0 - HALO + r, x0 - HALO + c);
    g2[r][c] = grey255(b, C, H, W, y0 - HALO + r, x0 - HALO + c);
  }
}

// sums[0] += sum dist * w, sums[1] += sum w,  w = valid * (1 - occ)
template <int M>
__global__ void __launch_bounds__(CT_THREADS)
census_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ img, const float* __restrict__ occ,
                  double* __restrict__ sums, int C, int H, int W) {
  constexpr int N = 2 * M + 1;
  __shared__ float gp[CT_H + 2 * M][CT_W + 2 * M], gi[CT_H + 2 * M][CT_W + 2 * M];
  const int b = blockIdx.z, x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
  const size_t HW = (size_t)H * W;
  stage_grey<M, M>(gp, gi, pred + (size_t)b * C * HW, img + (size_t)b * C * HW, C, H, W, y0, x0);
  __syncthreads();
  float acc[2] = {0.f, 0.f};
  const int tx = threadIdx.x % CT_W, ty0 = threadIdx.x / CT_W;
#pragma unroll
  for (int k = 0; k < CT_H / (CT_THREADS / CT_W); ++k) {
    const int ty = ty0 + k * (CT_THREADS / CT_W);
    const int y = y0 + ty, x = x0 + tx;
    if (y < M || y >= H - M || x < M || x >= W - M) continue;
    const float w = occ != nullptr ? 1.0f - __ldg(occ + (size_t)b * HW + (size_t)y * W + x) : 1.0f;
    const float cp = gp[ty + M][tx + M], ci = gi[ty + M][tx + M];
    float s = 0.f;
#pragma unroll
    for (int oy = 0; oy < N; ++oy)
#pragma unroll
      for (int ox = 0; ox < N; ++ox) {
        const float diff = soft_sign(gp[ty + oy][tx + ox] - cp) - soft_sign(gi[ty + oy][tx + ox] - ci);
        const float d = diff * diff;
        s += __fdividef(d, 0.1f + d);
      }
    acc[0] += s * (1.0f / (float)(N * N)) * w;
    acc[1] += w;
  }
  ocf_block_accumulate<2>(acc, sums);
}

// d_pred[c, q] = coef * 255 * k_c * ( sum_o w[q-o] G(q-o, o)  -  w[q] sum_o G(q, o) ),   G(p, o) = d softham / d u_pred(p, o)
template <int M>
__global__ void __launch_bounds__(CT_THREADS)
census_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ img, const float* __restrict__ occ,
                  const float* __restrict__ coef, float* __restrict__ d_pred, int C, int H, int W) {
  constexpr int N = 2 * M + 1, HALO = 2 * M;
  __shared__ float gp[CT_H + 2 * HALO][CT_W + 2 * HALO], gi[CT_H + 2 * HALO][CT_W + 2 * HALO];
  __shared__ float wt[CT_H + 2 * M][CT_W + 2 * M];  // weight of the patch centre p, halo M
  const int b = blockIdx.z, x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
  const size_t HW = (size_t)H * W;
  stage_grey<M, HALO>(gp, gi, pred + (size_t)b * C * HW, img + (size_t)b * C * HW, C, H, W, y0, x0);
  for (int i = threadIdx.x; i < (CT_W + 2 * M) * (CT_H + 2 * M); i += CT_THREADS) {
    const int r = i / (CT_W + 2 * M), c = i - r * (CT_W + 2 * M);
    const int y = y0 - M + r, x = x0 - M + c;
    float w = 0.f;
    if (y >= M && y < H - M && x >= M && x < W - M) w = occ != nullptr ? 1.0f - __ldg(occ + (size_t)b * HW + (size_t)y * W + x) : 1.0f;
    wt[r][c] = w;
  }
  __syncthreads();
  const float scale = __ldg(coef) * (255.0f / (float)(N * N));
  const int tx = threadIdx.x % CT_W, ty0 = threadIdx.x / CT_W;
  // G(p, o) with p at smem (py, px) [HALO-based coordinates] and neighbour p + o
  auto G = [&](int py, int px, int ny, int nx) -> float {
    const float up = gp[ny][nx] - gp[py][px], ui = gi[ny][nx] - gi[py][px];
    const float rp = rsqrtf(fmaf(up, up, 0.81f));
    const float diff = __fmul_rn(up, rp) - __fmul_rn(ui, rsqrtf(fmaf(ui, ui, 0.81f)));
    const float d = diff * diff;
    const float den = __fdividef(1.0f, 0.1f + d);
    // d/dt (t-ti)^2/(0.1+(t-ti)^2) = 0.2 diff / (0.1+d)^2 ;  dt/du = 0.81 (0.81+u^2)^-1.5
    return (0.2f * diff * den * den) * (0.81f * rp * rp * rp);
  };
#pragma unroll
  for (int k = 0; k < CT_H / (CT_THREADS / CT_W); ++k) {
    const int ty = ty0 + k * (CT_THREADS / CT_W);
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    const int qy = ty + HALO, qx = tx + HALO;
    float as_nb = 0.f, as_centre = 0.f;
#pragma unroll
    for (int oy = -M; oy <= M; ++oy)
#pragma unroll
      for (int ox = -M; ox <= M; ++ox) {
        // q is the neighbour (offset o) of the centre p = q - o
        const float wp = wt[ty + M - oy][tx + M - ox];
        if (wp != 0.f) as_nb = fmaf(wp, G(qy - oy, qx - ox, qy, qx), as_nb);
        as_centre += G(qy, qx, qy + oy, qx + ox);
      }
    const float gq = scale * (as_nb - wt[ty + M][tx + M] * as_centre);
    float* o = d_pred + (size_t)b * C * HW + (size_t)y * W + x;
    for (int c = 0; c < C; ++c) o[c * HW] = gq * grey_weight(C, c);
  }
}

}  // namespace

extern "C" int ocf_census_fwd(const float* pred, const float* img, const float* occ, double* sums, int B, int C, int H, int W,
                              int max_distance, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(pred); OCF_REQUIRE_PTR(img); OCF_REQUIRE_PTR(sums);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(max_distance >= 1 && max_distance <= 3, OCF_EUNSUPPORTED);
  OCF_REQUIRE(B <= 65535, OCF_EUNSUPPORTED);
  cudaStream_t s = ocf_cast_stream(stream);
  cudaError_t e = cudaMemsetAsync(sums, 0, 2 * sizeof(double), s);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((W + CT_W - 1) / CT_W, (H + CT_H - 1) / CT_H, B);
  if (max_distance == 1) census_fwd_kernel<1><<<grid, CT_THREADS, 0, s>>>(pred, img, occ, sums, C, H, W);
  else if (max_distance == 2) census_fwd_kernel<2><<<grid, CT_THREADS, 0, s>>>(pred, img, occ, sums, C, H, W);
  else census_fwd_kernel<3><<<grid, CT_THREADS, 0, s>>>(pred, img, occ, sums, C, H, W);
  return ocf_launch_status();
}

extern "C" int ocf_census_bwd(const float* pred, const float* img, const float* occ, const float* coef, float* d_pred, int B,
                              int C, int H, int W, int max_distance, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(pred); OCF_REQUIRE_PTR(img); OCF_REQUIRE_PTR(coef); OCF_REQUIRE_PTR(d_pred);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(max_distance >= 1 && max_distance <= 3, OCF_EUNSUPPORTED);
  OCF_REQUIRE(B <= 65535, OCF_EUNSUPPORTED);
  cudaStream_t s = ocf_cast_stream(stream);
  dim3 grid((W + CT_W - 1) / CT_W, (H + CT_H - 1) / CT_H, B);
  if (max_distance == 1) census_bwd_kernel<1><<<grid, CT_THREADS, 0, s>>>(pred, img, occ, coef, d_pred, C, H, W);
  else if (max_distance == 2) census_bwd_kernel<2><<<grid, CT_THREADS, 0, s>>>(pred, img, occ, coef, d_pred, C, H, W);
  else census_bwd_kernel<3><<<grid, CT_THREADS, 0, s>>>(pred, img, occ, coef, d_pred, C, H, W);
  return ocf_launch_status();
}
