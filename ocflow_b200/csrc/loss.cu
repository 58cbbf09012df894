// Occlusion-weighted photometric (Charbonnier), edge-aware smoothness and supervised losses, fp32.
//
// Replaces models/model.py:27-46 (robust_l1, photometric_error), :53-114 (gradient, first/second
// order smoothness), utils.py:8-18 (charbonnier_loss) and the loss assembly of
// general_step_occ_aware (models/model.py:379-407).  Every kernel is a single streaming pass: per-thread
// fp32 partials -> warp shuffle -> one double atomic per block, so the scalar losses are accumulated in
// fp64 (the reference sums in fp32 with ATen's pairwise tree; both sit far inside the 1e-3 loss bar).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int LT = 256;
// fused loss tuning knobs: pixels per thread (vector path) and grid cap in CTAs per SM
#ifndef OCF_OPF_VPX
#define OCF_OPF_VPX 4
#endif
#ifndef OCF_OPF_CAP
#define OCF_OPF_CAP 4
#endif

__device__ __forceinline__ float rho(float x, float a2) { return sqrtf(fmaf(x, x, a2)); }

__global__ void __launch_bounds__(LT) robust_l1_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n, float a2) {
  const size_t stride = (size_t)gridDim.x * LT;
  for (size_t i = (size_t)blockIdx.x * LT + threadIdx.x; i < n; i += stride) y[i] = rho(x[i], a2);
}

__global__ void __launch_bounds__(LT)
robust_l1_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x, float* __restrict__ gx, size_t n, float a2) {
  const size_t stride = (size_t)gridDim.x * LT;
  for (size_t i = (size_t)blockIdx.x * LT + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    gx[i] = gy[i] * v / rho(v, a2);
  }
}

// one thread per pixel (b, y, x); loops the C channels.  sums: [0] sum rho*(1-occ), [1] sum (1-occ)
__global__ void __launch_bounds__(LT)
photometric_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ img, const float* __restrict__ occ,
                       double* __restrict__ sums, int C, size_t HW, size_t npix, float a2) {
  float acc[2] = {0.f, 0.f};
  const size_t stride = (size_t)gridDim.x * LT;
  for (size_t p = (size_t)blockIdx.x * LT + threadIdx.x; p < npix; p += stride) {
    const size_t b = p / HW, q = p - b * HW;
    const float vis = occ != nullptr ? 1.0f - occ[p] : 1.0f;
    float e = 0.f;
    const size_t base = b * C * HW + q;
    for (int c = 0; c < C; ++c) e += rho(pred[base + c * HW] - img[base + c * HW], a2);
    acc[0] += e * vis;
    acc[1] += vis;
  }
  ocf_block_accumulate<2>(acc, sums);
}

__global__ void __launch_bounds__(LT)
photometric_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ img, const float* __restrict__ occ,
                       const float* __restrict__ coef, float* __restrict__ d_pred, float* __restrict__ d_img,
                       float* __restrict__ d_occ, int C, size_t HW, size_t npix, float a2) {
  const float k0 = coef[0], k1 = coef[1];
  const size_t stride = (size_t)gridDim.x * LT;
  for (size_t p = (size_t)blockIdx.x * LT + threadIdx.x; p < npix; p += stride) {
    const size_t b = p / HW, q = p - b * HW;
    const float vis = occ != nullptr ? 1.0f - occ[p] : 1.0f;
    const size_t base = b * C * HW + q;
    float e = 0.f;
    for (int c = 0; c < C; ++c) {
      const float d = pred[base + c * HW] - img[base + c * HW];
      const float r = rho(d, a2);
      e += r;
      const float gd = k0 * vis * d / r;
      if (d_pred != nullptr) d_pred[base + c * HW] = gd;
      if (d_img != nullptr) d_img[base + c * HW] = -gd;
    }
    // d/d occ of  k0*sum rho*(1-occ) + k1*sum(1-occ)
    if (d_occ != nullptr) d_occ[p] = -(k0 * e + k1);
  }
}

// ---- smoothness ---------------------------------------------------------------------------------
// order 1: w = exp(-mean_c (a*(I[p+1]-I[p]))^2), term = w * sum_cf rho(F[p+1]-F[p])            (model.py:93-101)
// order 2: w = exp(-mean_c (a*(I[p+2]-I[p]))^2), term = w * sum_cf rho((F[p+2]-F[p+1])-(F[p+1]-F[p]))  (:103-114)
// `step` is the element distance of one pixel along the direction (1 for x, W for y).
struct EdgeTerm {
  float w;   // edge weight
  float r;   // sum over flow channels of rho
};

__device__ __forceinline__ EdgeTerm edge_term(const float* __restrict__ ib, const float* __restrict__ fb, size_t q, size_t step,
                                              int Ci, int Cf, size_t HW, int order, float ae, float a2) {
  float s = 0.f;
  for (int c = 0; c < Ci; ++c) {
    const float d = ae * (ib[c * HW + q + order * step] - ib[c * HW + q]);
    s = fmaf(d, d, s);
  }
  EdgeTerm t;
  t.w = expf(-s / (float)Ci);
  t.r = 0.f;
  for (int c = 0; c < Cf; ++c) {
    const float* f = fb + c * HW + q;
    const float d = order == 1 ? f[step] - f[0] : (f[2 * step] - f[step]) - (f[step] - f[0]);
    t.r += rho(d, a2);
  }
  return t;
}

__global__ void __launch_bounds__(LT)
smooth_fwd_kernel(const float* __restrict__ img, const float* __restrict__ flow, double* __restrict__ sums, int Ci, int Cf, int H,
                  int W, size_t npix, int order, float ae, float a2) {
  const size_t HW = (size_t)H * W;
  float acc[2] = {0.f, 0.f};
  const size_t stride = (size_t)gridDim.x * LT;
  for (size_t p = (size_t)blockIdx.x * LT + threadIdx.x; p < npix; p += stride) {
    const size_t b = p / HW, q = p - b * HW;
    const int y = (int)(q / W), x = (int)(q - (size_t)y * W);
    const float* ib = img + b * Ci * HW;
    const float* fb = flow + b * Cf * HW;
    if (x + order < W) { const EdgeTerm t = edge_term(ib, fb, q, 1, Ci, Cf, HW, order, ae, a2); acc[0] += t.w * t.r; }
    if (y + order < H) { const EdgeTerm t = edge_term(ib, fb, q, W, Ci, Cf, HW, order, ae, a2); acc[1] += t.w * t.r; }
  }
  ocf_block_accumulate<2>(acc, sums);
}

// Backward as a scatter of each edge's contribution (atomics on d_flow / d_img, zeroed by the entry point).
__device__ __forceinline__ void edge_bwd(const float* __restrict__ ib, const float* __restrict__ fb, float* __restrict__ dib,
                                         float* __restrict__ dfb, size_t q, size_t step, int Ci, int Cf, size_t HW, int order,
                                         float ae, float a2, float k) {
  float s = 0.f;
  for (int c = 0; c < Ci; ++c) {
    const float d = ae * (ib[c * HW + q + order * step] - ib[c * HW + q]);
    s = fmaf(d, d, s);
  }
  const float w = expf(-s / (float)Ci);
  float r = 0.f;
  for (int c = 0; c < Cf; ++c) {
    const float* f = fb + c * HW + q;
    const float d = order == 1 ? f[step] - f[0] : (f[2 * step] - f[step]) - (f[step] - f[0]);
    const float rr = rho(d, a2);
    r += rr;
    if (dfb != nullptr) {
      const float t = k * w * d / rr;
      float* df = dfb + c * HW + q;
      if (order == 1) {
        atomicAdd(df + step, t);
        atomicAdd(df, -t);
      } else {
        atomicAdd(df + 2 * step, t);
        atomicAdd(df + step, -2.f * t);
        atomicAdd(df, t);
      }
    }
  }
  if (dib != nullptr) {
    // d w / d dI_c = w * (-2 a^2 / Ci) * dI_c
    const float kk = k * r * w * (-2.f * ae * ae / (float)Ci);
    for (int c = 0; c < Ci; ++c) {
      const float dI = ib[c * HW + q + order * step] - ib[c * HW + q];
      const float t = kk * dI;
      atomicAdd(dib + c * HW + q + order * step, t);
      atomicAdd(dib + c * HW + q, -t);
    }
  }
}

__global__ void __launch_bounds__(LT)
smooth_bwd_kernel(const float* __restrict__ img, const float* __restrict__ flow, const float* __restrict__ coef,
                  float* __restrict__ d_img, float* __restrict__ d_flow, int Ci, int Cf, int H, int W, size_t npix, int order,
                  float ae, float a2) {
  const size_t HW = (size_t)H * W;
  const float kx = coef[0], ky = coef[1];
  const size_t stride = (size_t)gridDim.x * LT;
  for (size_t p = (size_t)blockIdx.x * LT + threadIdx.x; p < npix; p += stride) {
    const size_t b = p / HW, q = p - b * HW;
    const int y = (int)(q / W), x = (int)(q - (size_t)y * W);
    const float* ib = img + b * Ci * HW;
    const float* fb = flow + b * Cf * HW;
    float* dib = d_img != nullptr ? d_img + b * Ci * HW : nullptr;
    float* dfb = d_flow != nullptr ? d_flow + b * Cf * HW : nullptr;
    if (x + order < W) edge_bwd(ib, fb, dib, dfb, q, 1, Ci, Cf, HW, order, ae, a2, kx);
    if (y + order < H) edge_bwd(ib, fb, dib, dfb, q, W, Ci, Cf, HW, order, ae, a2, ky);
  }
}

__global__ void __launch_bounds__(LT)
gradient_kernel(const float* __restrict__ img, float* __restrict__ dx, float* __restrict__ dy, int H, int W, int s, size_t n) {
  // n = B*C*H*W ; dx [.., H, W-s], dy [.., H-s, W]
  const size_t stride = (size_t)gridDim.x * LT;
  for (size_t i = (size_t)blockIdx.x * LT + threadIdx.x; i < n; i += stride) {
    const int x = (int)(i % W);
    const size_t row = i / W;
    const int y = (int)(row % H);
    const size_t plane = row / H;
    const float v = img[i];
    if (x + s < W) dx[(plane * H + y) * (size_t)(W - s) + x] = img[i + s] - v;
    if (y + s < H) dy[(plane * (size_t)(H - s) + y) * W + x] = img[i + (size_t)s * W] - v;
  }
}

// ---- fused occlusion-aware photometric pass -----------------------------------------------------
// A thread owns VPX consecutive pixels of one row (VPX = 4 when rows are 16-byte aligned: every regular stream --
// flow, range map, img1, flow_gt, occ_gt, d_flow -- moves as 128-bit accesses and the 4*VPX*C tap gathers of img2 are
// all in flight together; VPX = 1 is the ragged-width fallback).
template <int VPX>
struct PixVec {
  float v[VPX];
};

template <int VPX>
__device__ __forceinline__ PixVec<VPX> ld_vec(const float* p) {
  PixVec<VPX> r;
  if (VPX == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1 % VPX] = t.y; r.v[2 % VPX] = t.z; r.v[3 % VPX] = t.w;
  } else if (VPX == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    r.v[0] = t.x; r.v[1 % VPX] = t.y;
  } else {
#pragma unroll
    for (int i = 0; i < VPX; ++i) r.v[i] = p[i];
  }
  return r;
}

template <int VPX>
__device__ __forceinline__ void st_vec(float* p, const PixVec<VPX>& r) {
  if (VPX == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1 % VPX], r.v[2 % VPX], r.v[3 % VPX]);
  } else if (VPX == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(r.v[0], r.v[1 % VPX]);
  } else {
#pragma unroll
    for (int i = 0; i < VPX; ++i) p[i] = r.v[i];
  }
}

// grid (blocks over the pixel groups of ONE image, B): 32-bit index math only, no 64-bit divisions.
template <int VPX>
__global__ void __launch_bounds__(LT, 2)
occ_photo_fused_kernel(const float* __restrict__ img1, const float* __restrict__ img2, const float* __restrict__ flow,
                       const float* __restrict__ range, const float* __restrict__ flow_gt, const float* __restrict__ occ_gt,
                       double* __restrict__ sums, float* __restrict__ dflow, float* __restrict__ warped, int C, int H, int W,
                       float a2) {
  const int HW = H * W;
  const int gpi = HW / VPX;  // pixel groups per image
  const int b = blockIdx.y;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float fw = (float)(W - 1), fh = (float)(H - 1);
  const float dw = (float)max(W - 1, 1), dh = (float)max(H - 1, 1);
  const float cx = (0.5f * fw) * (2.0f / dw), cy = (0.5f * fh) * (2.0f / dh);  // align_corners=True chain factor (== 1)
  const float* fl_b = flow + (size_t)b * 2 * HW;
  const float* i1_b = img1 + (size_t)b * C * HW;
  const float* i2_b = img2 + (size_t)b * C * HW;
  for (int gi = blockIdx.x * LT + threadIdx.x; gi < gpi; gi += gridDim.x * LT) {
    const int q = gi * VPX;
    const int y = q / W, x = q - y * W;
    const PixVec<VPX> U = ld_vec<VPX>(fl_b + q), V = ld_vec<VPX>(fl_b + HW + q);
    int onw[VPX], one[VPX], osw[VPX], ose[VPX];
    float wx0[VPX], wx1[VPX], wy0[VPX], wy1[VPX];
    bool vnw[VPX], vne[VPX], vsw[VPX], vse[VPX];
#pragma unroll
    for (int i = 0; i < VPX; ++i) {
      // align_corners=True coordinates, reference op order (models/model.py:211-212 + ATen unnormalize); x*0.5f == x/2 exactly
      float ix = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)(x + i), U.v[i])), dw), 1.0f), 1.0f), 0.5f), fw);
      float iy = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)y, V.v[i])), dh), 1.0f), 1.0f), 0.5f), fh);
      if (!(ix > -2147483648.0f && ix < 2147483520.0f)) ix = -100.f;
      if (!(iy > -2147483648.0f && iy < 2147483520.0f)) iy = -100.f;
      const float fx0 = floorf(ix), fy0 = floorf(iy);
      const int x0 = (int)fx0, y0 = (int)fy0;
      wx1[i] = ix - fx0; wx0[i] = (fx0 + 1.f) - ix; wy1[i] = iy - fy0; wy0[i] = (fy0 + 1.f) - iy;
      const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
      const bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
      vnw[i] = vx0 && vy0; vne[i] = vx1 && vy0; vsw[i] = vx0 && vy1; vse[i] = vx1 && vy1;
      const int o = y0 * W + x0;
      onw[i] = vnw[i] ? o : 0; one[i] = vne[i] ? o + 1 : 0; osw[i] = vsw[i] ? o + W : 0; ose[i] = vse[i] ? o + W + 1 : 0;
    }
    PixVec<VPX> occv, vis;
    if (range != nullptr) {
      const PixVec<VPX> r = ld_vec<VPX>(range + (size_t)b * HW + q);
#pragma unroll
      for (int i = 0; i < VPX; ++i) occv.v[i] = 1.0f - fminf(fmaxf(r.v[i], 0.f), 1.f);
    } else {
#pragma unroll
      for (int i = 0; i < VPX; ++i) occv.v[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < VPX; ++i) vis.v[i] = 1.0f - occv.v[i];
    float e[VPX], gx[VPX], gy[VPX];
#pragma unroll
    for (int i = 0; i < VPX; ++i) { e[i] = 0.f; gx[i] = 0.f; gy[i] = 0.f; }
    for (int c = 0; c < C; ++c) {
      const float* ip = i2_b + c * HW;
      const PixVec<VPX> t1 = ld_vec<VPX>(i1_b + c * HW + q);
      float a[VPX], bb[VPX], cc[VPX], dd[VPX];
#pragma unroll
      for (int i = 0; i < VPX; ++i) {
        a[i] = vnw[i] ? __ldg(ip + onw[i]) : 0.f; bb[i] = vne[i] ? __ldg(ip + one[i]) : 0.f;
        cc[i] = vsw[i] ? __ldg(ip + osw[i]) : 0.f; dd[i] = vse[i] ? __ldg(ip + ose[i]) : 0.f;
      }
      PixVec<VPX> wv;
#pragma unroll
      for (int i = 0; i < VPX; ++i) {
        float s = 0.f;
        s = fmaf(a[i], wx0[i] * wy0[i], s); s = fmaf(bb[i], wx1[i] * wy0[i], s);
        s = fmaf(cc[i], wx0[i] * wy1[i], s); s = fmaf(dd[i], wx1[i] * wy1[i], s);
        wv.v[i] = s;
        const float d = s - t1.v[i];
        const float r2 = fmaf(d, d, a2);
        const float inv = rsqrtf(r2);     // rho = r2 * inv, rho' = d * inv   (MUFU.RSQ, <= 2 ulp: far inside the loss tolerance)
        e[i] = fmaf(r2, inv, e[i]);
        const float gr = d * inv;
        gx[i] = fmaf(gr, (bb[i] - a[i]) * wy0[i] + (dd[i] - cc[i]) * wy1[i], gx[i]);
        gy[i] = fmaf(gr, (cc[i] - a[i]) * wx0[i] + (dd[i] - bb[i]) * wx1[i], gy[i]);
      }
      if (warped != nullptr) st_vec<VPX>(warped + ((size_t)b * C + c) * HW + q, wv);
    }
#pragma unroll
    for (int i = 0; i < VPX; ++i) {
      acc[0] += e[i] * vis.v[i]; acc[1] += vis.v[i]; acc[2] += e[i] * occv.v[i]; acc[3] += occv.v[i];
    }
    if (dflow != nullptr) {
      PixVec<VPX> ox, oy;
#pragma unroll
      for (int i = 0; i < VPX; ++i) { ox.v[i] = gx[i] * vis.v[i] * cx; oy.v[i] = gy[i] * vis.v[i] * cy; }
      st_vec<VPX>(dflow + (size_t)b * 2 * HW + q, ox);
      st_vec<VPX>(dflow + (size_t)b * 2 * HW + HW + q, oy);
    }
    if (flow_gt != nullptr) {
      const PixVec<VPX> gu = ld_vec<VPX>(flow_gt + (size_t)b * 2 * HW + q), gv = ld_vec<VPX>(flow_gt + (size_t)b * 2 * HW + HW + q);
#pragma unroll
      for (int i = 0; i < VPX; ++i) {
        const float du = U.v[i] - gu.v[i], dv = V.v[i] - gv.v[i];
        acc[4] += du * du + dv * dv;
      }
    }
    if (occ_gt != nullptr) {
      // F.binary_cross_entropy(input=occ_gt, target=occ_pred)  -- swapped on purpose, models/model.py:407
      const PixVec<VPX> pin = ld_vec<VPX>(occ_gt + (size_t)b * HW + q);
#pragma unroll
      for (int i = 0; i < VPX; ++i) {
        const float lp = fmaxf(logf(pin.v[i]), -100.f), l1p = fmaxf(logf(1.0f - pin.v[i]), -100.f);
        acc[5] += -(occv.v[i] * lp + (1.0f - occv.v[i]) * l1p);
      }
    }
  }
  ocf_block_accumulate<6>(acc, sums);
}

// Quad version (W % 4 == 0, 16-byte aligned tensors, C == 3): a thread owns 4 horizontally adjacent pixels.  The kernel is
// LATENCY bound (a pixel needs flow -> coordinates -> gathers -> loss: dependent DRAM round trips), so everything that does
// not depend on the flow -- img1 (3 channels), range map, flow_gt, occ_gt -- is requested together with the flow, and the
// gathers of all three channels are issued back to back: two exposed round trips per quad instead of six.  When the 4
// bilinear samples are coherent (one tap-row pair, consecutive columns -- the normal case for a network flow) the 16 tap
// gathers per channel are 4 x LDG.128 (two aligned groups per tap row) + a register funnel (common.cuh); otherwise the 16
// scalar gathers are issued.
template <int MINB>
__global__ void __launch_bounds__(LT, MINB)
occ_photo_quad_kernel(const float* __restrict__ img1, const float* __restrict__ img2, const float* __restrict__ flow,
                      const float* __restrict__ range, const float* __restrict__ flow_gt, const float* __restrict__ occ_gt,
                      double* __restrict__ sums, float* __restrict__ dflow, float* __restrict__ warped, int H, int W, float a2) {
  constexpr int C = 3;
  const int HW = H * W;
  const int gpi = HW >> 2;  // quads per image
  const int b = blockIdx.y;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float fw = (float)(W - 1), fh = (float)(H - 1);
  const float dw = (float)max(W - 1, 1), dh = (float)max(H - 1, 1);
  const float cx = (0.5f * fw) * (2.0f / dw), cy = (0.5f * fh) * (2.0f / dh);  // align_corners=True chain factor (== 1)
  const float* fl_b = flow + (size_t)b * 2 * HW;
  const float* i1_b = img1 + (size_t)b * C * HW;
  const float* i2_b = img2 + (size_t)b * C * HW;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int gi = blockIdx.x * LT + threadIdx.x; gi < gpi; gi += gridDim.x * LT) {
    const int q = gi << 2;
    const int y = q / W, x = q - y * W;
    // ---- round trip 1: the flow and everything that does not depend on it ----
    const float4 U4 = __ldg(reinterpret_cast<const float4*>(fl_b + q)), V4 = __ldg(reinterpret_cast<const float4*>(fl_b + HW + q));
    float4 T1[C];
#pragma unroll
    for (int c = 0; c < C; ++c) T1[c] = ocf_ldg_stream4(i1_b + c * HW + q);
    const float4 r4 = range != nullptr ? ocf_ldg_stream4(range + (size_t)b * HW + q) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 gu = flow_gt != nullptr ? ocf_ldg_stream4(flow_gt + (size_t)b * 2 * HW + q) : zero4;
    const float4 gv = flow_gt != nullptr ? ocf_ldg_stream4(flow_gt + (size_t)b * 2 * HW + HW + q) : zero4;
    const float4 p4 = occ_gt != nullptr ? ocf_ldg_stream4(occ_gt + (size_t)b * HW + q) : zero4;
    const float U[4] = {U4.x, U4.y, U4.z, U4.w}, V[4] = {V4.x, V4.y, V4.z, V4.w};
    int x0[4], y0[4];
    float wx0[4], wx1[4], wy0[4], wy1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // align_corners=True coordinates, reference op order (models/model.py:211-212 + ATen unnormalize); x*0.5f == x/2 exactly
      float ix = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)(x + i), U[i])), dw), 1.0f), 1.0f), 0.5f), fw);
      float iy = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)y, V[i])), dh), 1.0f), 1.0f), 0.5f), fh);
      if (!(ix > -2147483648.0f && ix < 2147483520.0f)) ix = -100.f;
      if (!(iy > -2147483648.0f && iy < 2147483520.0f)) iy = -100.f;
      const float fx0 = floorf(ix), fy0 = floorf(iy);
      x0[i] = (int)fx0; y0[i] = (int)fy0;
      wx1[i] = ix - fx0; wx0[i] = (fx0 + 1.f) - ix; wy1[i] = iy - fy0; wy0[i] = (fy0 + 1.f) - iy;
    }
    bool fast = x0[0] >= -8 && x0[0] <= W && y0[0] >= -2 && y0[0] <= H;
#pragma unroll
    for (int i = 1; i < 4; ++i) fast = fast && y0[i] == y0[0] && x0[i] == x0[0] + i;
    // ---- round trip 2: the gathers of all channels ----
    float ta[C][4], tb[C][4], tc[C][4], td[C][4];
    if (fast) {
      const int a = x0[0] & ~3, o = x0[0] - a;
      const bool vn = y0[0] >= 0 && y0[0] < H, vs = y0[0] + 1 >= 0 && y0[0] + 1 < H;
      const bool va = a >= 0 && a + 3 < W, vb = a + 4 >= 0 && a + 7 < W;
      const float* rp = i2_b + y0[0] * W + a;
      float4 a0[C], a1[C], b0[C], b1[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        a0[c] = ldg4_or_zero(rp + c * HW, vn && va); a1[c] = ldg4_or_zero(rp + c * HW + 4, vn && vb);
        b0[c] = ldg4_or_zero(rp + c * HW + W, vs && va); b1[c] = ldg4_or_zero(rp + c * HW + W + 4, vs && vb);
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float qn[8] = {a0[c].x, a0[c].y, a0[c].z, a0[c].w, a1[c].x, a1[c].y, a1[c].z, a1[c].w};
        const float qs[8] = {b0[c].x, b0[c].y, b0[c].z, b0[c].w, b1[c].x, b1[c].y, b1[c].z, b1[c].w};
        float n5[5], s5[5];
        funnel_gather(qn, o, n5);
        funnel_gather(qs, o, s5);
#pragma unroll
        for (int i = 0; i < 4; ++i) { ta[c][i] = n5[i]; tb[c][i] = n5[i + 1]; tc[c][i] = s5[i]; td[c][i] = s5[i + 1]; }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool vx0 = x0[i] >= 0 && x0[i] < W, vx1 = x0[i] + 1 >= 0 && x0[i] + 1 < W;
        const bool vy0 = y0[i] >= 0 && y0[i] < H, vy1 = y0[i] + 1 >= 0 && y0[i] + 1 < H;
        const int off = y0[i] * W + x0[i];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float* ip = i2_b + c * HW;
          ta[c][i] = (vx0 && vy0) ? __ldg(ip + off) : 0.f; tb[c][i] = (vx1 && vy0) ? __ldg(ip + off + 1) : 0.f;
          tc[c][i] = (vx0 && vy1) ? __ldg(ip + off + W) : 0.f; td[c][i] = (vx1 && vy1) ? __ldg(ip + off + W + 1) : 0.f;
        }
      }
    }
    float e[4] = {0.f, 0.f, 0.f, 0.f}, gx[4] = {0.f, 0.f, 0.f, 0.f}, gy[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float t1[4] = {T1[c].x, T1[c].y, T1[c].z, T1[c].w};
      float wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float sacc = 0.f;
        sacc = fmaf(ta[c][i], wx0[i] * wy0[i], sacc); sacc = fmaf(tb[c][i], wx1[i] * wy0[i], sacc);
        sacc = fmaf(tc[c][i], wx0[i] * wy1[i], sacc); sacc = fmaf(td[c][i], wx1[i] * wy1[i], sacc);
        wv[i] = sacc;
        const float d = sacc - t1[i];
        const float r2 = fmaf(d, d, a2);
        const float inv = rsqrtf(r2);     // rho = r2 * inv, rho' = d * inv   (MUFU.RSQ, <= 2 ulp: far inside the loss tolerance)
        e[i] = fmaf(r2, inv, e[i]);
        const float gr = d * inv;
        gx[i] = fmaf(gr, (tb[c][i] - ta[c][i]) * wy0[i] + (td[c][i] - tc[c][i]) * wy1[i], gx[i]);
        gy[i] = fmaf(gr, (tc[c][i] - ta[c][i]) * wx0[i] + (td[c][i] - tb[c][i]) * wx1[i], gy[i]);
      }
      if (warped != nullptr) *reinterpret_cast<float4*>(warped + ((size_t)b * C + c) * HW + q) = make_float4(wv[0], wv[1], wv[2], wv[3]);
    }
    const float r[4] = {r4.x, r4.y, r4.z, r4.w};
    float occv[4], vis[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      occv[i] = range != nullptr ? 1.0f - fminf(fmaxf(r[i], 0.f), 1.f) : 0.f;
      vis[i] = 1.0f - occv[i];
      acc[0] += e[i] * vis[i]; acc[1] += vis[i]; acc[2] += e[i] * occv[i]; acc[3] += occv[i];
    }
    if (dflow != nullptr) {
      float* df = dflow + (size_t)b * 2 * HW + q;
      *reinterpret_cast<float4*>(df) = make_float4(gx[0] * vis[0] * cx, gx[1] * vis[1] * cx, gx[2] * vis[2] * cx, gx[3] * vis[3] * cx);
      *reinterpret_cast<float4*>(df + HW) = make_float4(gy[0] * vis[0] * cy, gy[1] * vis[1] * cy, gy[2] * vis[2] * cy, gy[3] * vis[3] * cy);
    }
    if (flow_gt != nullptr) {
      const float gus[4] = {gu.x, gu.y, gu.z, gu.w}, gvs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float du = U[i] - gus[i], dv = V[i] - gvs[i];
        acc[4] += du * du + dv * dv;
      }
    }
    if (occ_gt != nullptr) {
      // F.binary_cross_entropy(input=occ_gt, target=occ_pred)  -- swapped on purpose, models/model.py:407
      const float pin[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float lp = fmaxf(logf(pin[i]), -100.f), l1p = fmaxf(logf(1.0f - pin[i]), -100.f);
        acc[5] += -(occv[i] * lp + (1.0f - occv[i]) * l1p);
      }
    }
  }
  ocf_block_accumulate<6>(acc, sums);
}

// ---- supervised pair losses ---------------------------------------------------------------------
__global__ void __launch_bounds__(LT)
pair_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ sum, float* __restrict__ grad,
                 size_t n, int kind) {
  float acc[1] = {0.f};
  const size_t stride = (size_t)gridDim.x * LT;
  for (size_t i = (size_t)blockIdx.x * LT + threadIdx.x; i < n; i += stride) {
    const float p = a[i], t = b[i];
    float l, g;
    if (kind == 0) {        // L1
      const float d = p - t;
      l = fabsf(d);
      g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    } else if (kind == 1) { // squared error
      const float d = p - t;
      l = d * d;
      g = 2.f * d;
    } else {                // BCE(p, t) with ATen's -100 clamp on the logs
      const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.0f - p), -100.f);
      const float bce = -(t * lp + (1.0f - t) * l1p);
      // ATen binary_cross_entropy_backward: (p - t) / max((1-p)*p, 1e-12)
      const float dbce = (p - t) / fmaxf((1.0f - p) * p, 1e-12f);
      if (kind == 2) { l = bce; g = dbce; }
      else {                // focal, gamma = 2: (1-exp(-bce))^2 * bce     (occlusion_model.py:55-62)
        const float pt = expf(-bce), om = 1.0f - pt;
        l = om * om * bce;
        g = (2.f * om * pt * bce + om * om) * dbce;
      }
    }
    acc[0] += l;
    if (grad != nullptr) grad[i] = g;
  }
  ocf_block_accumulate<1>(acc, sum);
}

inline unsigned stream_grid(size_t n) {
  size_t blocks = (n + LT - 1) / LT;
  const size_t cap = (size_t)OCF_SM_COUNT * 8;  // 8 resident 256-thread CTAs per SM
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks ? blocks : 1);
}

}  // namespace

extern "C" int ocf_robust_l1_fwd(const float* x, float* y, long long n, float alpha, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(x); OCF_REQUIRE_PTR(y);
  OCF_REQUIRE(n > 0, OCF_ESHAPE);
  robust_l1_fwd_kernel<<<stream_grid((size_t)n), LT, 0, ocf_cast_stream(stream)>>>(x, y, (size_t)n, alpha * alpha);
  return ocf_launch_status();
}

extern "C" int ocf_robust_l1_bwd(const float* grad_y, const float* x, float* grad_x, long long n, float alpha, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(grad_y); OCF_REQUIRE_PTR(x); OCF_REQUIRE_PTR(grad_x);
  OCF_REQUIRE(n > 0, OCF_ESHAPE);
  robust_l1_bwd_kernel<<<stream_grid((size_t)n), LT, 0, ocf_cast_stream(stream)>>>(grad_y, x, grad_x, (size_t)n, alpha * alpha);
  return ocf_launch_status();
}

extern "C" int ocf_photometric_fwd(const float* pred, const float* img, const float* occ, double* sums, int B, int C, int H, int W,
                                   float alpha, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(pred); OCF_REQUIRE_PTR(img); OCF_REQUIRE_PTR(sums);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  cudaStream_t s = ocf_cast_stream(stream);
  cudaError_t e = cudaMemsetAsync(sums, 0, 2 * sizeof(double), s);
  if (e != cudaSuccess) return (int)e;
  const size_t HW = (size_t)H * W, npix = HW * B;
  photometric_fwd_kernel<<<stream_grid(npix), LT, 0, s>>>(pred, img, occ, sums, C, HW, npix, alpha * alpha);
  return ocf_launch_status();
}

extern "C" int ocf_photometric_bwd(const float* pred, const float* img, const float* occ, const float* coef, float* d_pred,
                                   float* d_img, float* d_occ, int B, int C, int H, int W, float alpha, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(pred); OCF_REQUIRE_PTR(img); OCF_REQUIRE_PTR(coef);
  OCF_REQUIRE(d_pred != nullptr || d_img != nullptr || d_occ != nullptr, OCF_ENULL);
  OCF_REQUIRE(d_occ == nullptr || occ != nullptr, OCF_ENULL);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  const size_t HW = (size_t)H * W, npix = HW * B;
  photometric_bwd_kernel<<<stream_grid(npix), LT, 0, ocf_cast_stream(stream)>>>(pred, img, occ, coef, d_pred, d_img, d_occ, C, HW,
                                                                                npix, alpha * alpha);
  return ocf_launch_status();
}

extern "C" int ocf_smooth_fwd(const float* img, const float* flow, double* sums, int B, int Ci, int Cf, int H, int W, int order,
                              float alpha_edge, float alpha_rho, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(img); OCF_REQUIRE_PTR(flow); OCF_REQUIRE_PTR(sums);
  OCF_REQUIRE(B > 0 && Ci > 0 && Cf > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(order == 1 || order == 2, OCF_EUNSUPPORTED);
  cudaStream_t s = ocf_cast_stream(stream);
  cudaError_t e = cudaMemsetAsync(sums, 0, 2 * sizeof(double), s);
  if (e != cudaSuccess) return (int)e;
  const size_t npix = (size_t)B * H * W;
  smooth_fwd_kernel<<<stream_grid(npix), LT, 0, s>>>(img, flow, sums, Ci, Cf, H, W, npix, order, alpha_edge, alpha_rho * alpha_rho);
  return ocf_launch_status();
}

extern "C" int ocf_smooth_bwd(const float* img, const float* flow, const float* coef, float* d_img, float* d_flow, int B, int Ci,
                              int Cf, int H, int W, int order, float alpha_edge, float alpha_rho, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(img); OCF_REQUIRE_PTR(flow); OCF_REQUIRE_PTR(coef);
  OCF_REQUIRE(d_img != nullptr || d_flow != nullptr, OCF_ENULL);
  OCF_REQUIRE(B > 0 && Ci > 0 && Cf > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(order == 1 || order == 2, OCF_EUNSUPPORTED);
  cudaStream_t s = ocf_cast_stream(stream);
  const size_t npix = (size_t)B * H * W;
  cudaError_t e;
  if (d_img != nullptr && (e = cudaMemsetAsync(d_img, 0, sizeof(float) * npix * Ci, s)) != cudaSuccess) return (int)e;
  if (d_flow != nullptr && (e = cudaMemsetAsync(d_flow, 0, sizeof(float) * npix * Cf, s)) != cudaSuccess) return (int)e;
  smooth_bwd_kernel<<<stream_grid(npix), LT, 0, s>>>(img, flow, coef, d_img, d_flow, Ci, Cf, H, W, npix, order, alpha_edge,
                                                     alpha_rho * alpha_rho);
  return ocf_launch_status();
}

extern "C" int ocf_gradient(const float* img, float* dx, float* dy, int B, int C, int H, int W, int stride, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(img); OCF_REQUIRE_PTR(dx); OCF_REQUIRE_PTR(dy);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(stride >= 1 && stride < H && stride < W, OCF_ESHAPE);
  const size_t n = (size_t)B * C * H * W;
  gradient_kernel<<<stream_grid(n), LT, 0, ocf_cast_stream(stream)>>>(img, dx, dy, H, W, stride, n);
  return ocf_launch_status();
}

extern "C" int ocf_occ_photo_fused(const float* img1, const float* img2, const float* flow, const float* range_map,
                                   const float* flow_gt, const float* occ_gt, double* sums, float* dflow_unit, float* warped_out,
                                   int B, int C, int H, int W, float alpha, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(img1); OCF_REQUIRE_PTR(img2); OCF_REQUIRE_PTR(flow); OCF_REQUIRE_PTR(sums);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  cudaStream_t s = ocf_cast_stream(stream);
  cudaError_t e = cudaMemsetAsync(sums, 0, 8 * sizeof(double), s);
  if (e != cudaSuccess) return (int)e;
  bool vec = (W % 4 == 0) && ocf_aligned16(img1) && ocf_aligned16(flow);
  const float* opt[5] = {range_map, flow_gt, occ_gt, dflow_unit, warped_out};
  for (const float* p : opt) vec = vec && (p == nullptr || ocf_aligned16(p));
  OCF_REQUIRE((long long)C * H * W < (1LL << 31) && B <= 65535, OCF_EUNSUPPORTED);
  const int vpx = vec ? OCF_OPF_VPX : 1;
  const int gpi = H * W / vpx;
  // blocks per image: every thread owns at most a few pixel groups; the grid is capped at OCF_OPF_CAP CTAs per SM
  int bx = (gpi + LT - 1) / LT;
  const int cap = (OCF_OPF_CAP * OCF_SM_COUNT + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(bx, B);
  if (vec && OCF_OPF_VPX == 4 && C == 3) {
    // one quad per thread (no grid-stride repetition needed up to the cap): enough CTAs for >= 6 per SM
    int qx = (gpi + LT - 1) / LT;
    static const int qcapn = []() { const char* e = getenv("OCF_OPF_QCAP"); return e ? atoi(e) : 8; }();
    const int qcap = (qcapn * OCF_SM_COUNT + B - 1) / B;
    if (qx > qcap) qx = qcap;
    static const int minb = []() { const char* e = getenv("OCF_OPF_MINB"); return e ? atoi(e) : 2; }();   // developer knob (tuning runs)
    if (minb == 3)
      occ_photo_quad_kernel<3><<<dim3(qx, B), LT, 0, s>>>(img1, img2, flow, range_map, flow_gt, occ_gt, sums, dflow_unit, warped_out, H, W, alpha * alpha);
    else
      occ_photo_quad_kernel<2><<<dim3(qx, B), LT, 0, s>>>(img1, img2, flow, range_map, flow_gt, occ_gt, sums, dflow_unit, warped_out, H, W, alpha * alpha);
  } else if (vec)
    occ_photo_fused_kernel<OCF_OPF_VPX><<<grid, LT, 0, s>>>(img1, img2, flow, range_map, flow_gt, occ_gt, sums, dflow_unit, warped_out, C, H, W, alpha * alpha);
  else
    occ_photo_fused_kernel<1><<<grid, LT, 0, s>>>(img1, img2, flow, range_map, flow_gt, occ_gt, sums, dflow_unit, warped_out, C, H, W, alpha * alpha);
  return ocf_launch_status();
}

extern "C" int ocf_pair_loss(const float* a, const float* b, double* sum_out, float* grad, long long n, int kind, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(a); OCF_REQUIRE_PTR(b); OCF_REQUIRE_PTR(sum_out);
  OCF_REQUIRE(n > 0, OCF_ESHAPE);
  OCF_REQUIRE(kind >= 0 && kind <= 3, OCF_EUNSUPPORTED);
  cudaStream_t s = ocf_cast_stream(stream);
  cudaError_t e = cudaMemsetAsync(sum_out, 0, sizeof(double), s);
  if (e != cudaSuccess) return (int)e;
  pair_loss_kernel<<<stream_grid((size_t)n), LT, 0, s>>>(a, b, sum_out, grad, (size_t)n, kind);
  return ocf_launch_status();
}
