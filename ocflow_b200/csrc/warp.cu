// Bilinear backward warp (forward gather, backward scatter) and range-map forward splat, fp32 NCHW.
//
// Replaces the reference's 11 `warp` / `backwarp` bodies (meshgrid on the CPU + H2D copy + ~15 ATen
// launches + F.grid_sample, e.g. utils.py:20-58, cost_volume_flow_net.py:121-151) and
// compute_range_map (models/model.py:243-305: 4 x nonzero host syncs + scatter_add_).
//
// Coordinates follow the reference + ATen op order exactly, with FMA contraction suppressed:
//   g  = 2*(x+u)/max(W-1,1) - 1                      (utils.py:43-44)
//   ix = ((g+1)/2)*(W-1)          align_corners=True  (ATen GridSampler.h grid_sampler_unnormalize)
//   ix = ((g+1)*W-1)/2            align_corners=False
// so align_corners=False samples at (x+u)*W/(W-1) - 0.5, the reference's quirk.
#include <stdlib.h>

#include "common.cuh"

namespace {

struct Taps {
  int x0, y0;          // north-west tap
  float wx0, wx1, wy0, wy1;
  bool vx0, vx1, vy0, vy1;
};

__device__ __forceinline__ float unnormalize(float v, int size, int denom, bool align) {
  // v: pixel + flow.  reference normalisation then ATen un-normalisation, one rounding per op.
  float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, v), (float)denom), 1.0f);
  float r;
  if (align) r = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.0f), 2.0f), (float)(size - 1));
  else r = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.0f), (float)size), 1.0f), 2.0f);
  // ATen compute_coordinates -> safe_downgrade_to_int_range
  if (!(r > -2147483648.0f && r < 2147483520.0f)) r = -100.0f;
  return r;
}

__device__ __forceinline__ Taps make_taps(float ix, float iy, int H, int W) {
  Taps t;
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  t.x0 = (int)fx0;
  t.y0 = (int)fy0;
  t.wx1 = ix - fx0;            // weight of the east taps
  t.wx0 = (fx0 + 1.0f) - ix;   // weight of the west taps  (ATen: ix_se - ix)
  t.wy1 = iy - fy0;
  t.wy0 = (fy0 + 1.0f) - iy;
  t.vx0 = t.x0 >= 0 && t.x0 < W;
  t.vx1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  t.vy0 = t.y0 >= 0 && t.y0 < H;
  t.vy1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  return t;
}

__device__ __forceinline__ float cover_of(const Taps& t) {
  float m = 0.f;
  if (t.vx0 && t.vy0) m += t.wx0 * t.wy0;
  if (t.vx1 && t.vy0) m += t.wx1 * t.wy0;
  if (t.vx0 && t.vy1) m += t.wx0 * t.wy1;
  if (t.vx1 && t.vy1) m += t.wx1 * t.wy1;
  return m;
}

// utils.py:54-55: mask<0.9999 -> 0 ; mask>0 -> 1
__device__ __forceinline__ float mask_of(const Taps& t) { return cover_of(t) < 0.9999f ? 0.f : 1.f; }

constexpr int WARP_THREADS = 256;
// warp backward tuning knobs: rows per thread and channels whose loads are batched ahead of the atomics
#ifndef OCF_WB_R
#define OCF_WB_R 1
#endif
#ifndef OCF_WB_CB
#define OCF_WB_CB 2
#endif

// grid: (ceil(HW/256), channel slabs, B)
__global__ void __launch_bounds__(WARP_THREADS)
warp_fwd_kernel(const float* __restrict__ img, const float* __restrict__ flow, const float* __restrict__ occ,
                float* __restrict__ out, int C, int H, int W, int slab, int flags, float scale) {
  const int pix = blockIdx.x * WARP_THREADS + threadIdx.x;
  const int HW = H * W;
  if (pix >= HW) return;
  const int y = pix / W, x = pix - y * W;
  const int b = blockIdx.z;
  const bool align = flags & OCF_WARP_ALIGN_CORNERS;
  const float u = __fmul_rn(__ldg(flow + ((size_t)b * 2) * HW + pix), scale);
  const float v = __fmul_rn(__ldg(flow + ((size_t)b * 2 + 1) * HW + pix), scale);
  const float ix = unnormalize(__fadd_rn((float)x, u), W, max(W - 1, 1), align);
  const float iy = unnormalize(__fadd_rn((float)y, v), H, max(H - 1, 1), align);
  const Taps t = make_taps(ix, iy, H, W);
  float mul = 1.f;
  if (flags & OCF_WARP_IS_MASK) mul = mask_of(t);
  if (occ != nullptr) mul *= __ldg(occ + (size_t)b * HW + pix);
  const float wnw = t.wx0 * t.wy0, wne = t.wx1 * t.wy0, wsw = t.wx0 * t.wy1, wse = t.wx1 * t.wy1;
  const bool vnw = t.vx0 && t.vy0, vne = t.vx1 && t.vy0, vsw = t.vx0 && t.vy1, vse = t.vx1 && t.vy1;
  // clamp the tap offsets so that even dropped taps form a valid address (never dereferenced)
  const int onw = vnw ? t.y0 * W + t.x0 : 0, one = vne ? t.y0 * W + t.x0 + 1 : 0;
  const int osw = vsw ? (t.y0 + 1) * W + t.x0 : 0, ose = vse ? (t.y0 + 1) * W + t.x0 + 1 : 0;
  const int c_begin = blockIdx.y * slab, c_end = min(C, c_begin + slab);
  const float* ip = img + ((size_t)b * C + c_begin) * HW;
  float* op = out + ((size_t)b * C + c_begin) * HW + pix;
#pragma unroll 4
  for (int c = c_begin; c < c_end; ++c, ip += HW, op += HW) {
    float r = 0.f;
    if (vnw) r = fmaf(__ldg(ip + onw), wnw, r);
    if (vne) r = fmaf(__ldg(ip + one), wne, r);
    if (vsw) r = fmaf(__ldg(ip + osw), wsw, r);
    if (vse) r = fmaf(__ldg(ip + ose), wse, r);
    *op = r * mul;
  }
}

// Backward.  d_img is a scatter (red.global.add.f32, zeroed by the entry point); d_flow / d_occ are per-pixel and
// accumulated across channel slabs with one atomic per slab (plain store when there is 1 slab).
// Warp-aggregated scatter: a thread owns R vertically adjacent pixels and a warp 32 horizontally adjacent columns.
// For a smooth flow the south taps of row r are the north taps of row r+1 and the east taps of lane l are the west
// taps of lane l+1; both coincidences are detected once per pixel (they do not depend on the channel) and the
// contributions are summed in registers / with one shuffle before the atomic, so a fully coherent neighbourhood
// issues (R+1)/R atomics per pixel and channel instead of 4.
// (Measured, profiles/: every lane of a scalar red is its own 32-byte sector-op in the L2 atomic unit; a variant that
// re-aligned coherent 8-lane segments into red.global.add.v4.f32 cut sector-ops by a third but not the run time -- the
// kernel is bound by issue slots and load latency, not by the atomic unit -- so the simpler scalar form is kept.)
template <int R>
__global__ void __launch_bounds__(WARP_THREADS, 2)
warp_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ img, const float* __restrict__ flow,
                const float* __restrict__ occ, float* __restrict__ d_img, float* __restrict__ d_flow,
                float* __restrict__ d_occ, int C, int H, int W, int slab, int nslabs, int flags, float scale) {
  const int HW = H * W;
  const int HG = (H + R - 1) / R;  // row groups
  const int gidx = blockIdx.x * WARP_THREADS + threadIdx.x;
  const bool in_grid = gidx < HG * W;
  const int gc = in_grid ? gidx : HG * W - 1;
  const int yg = gc / W, x = gc - yg * W;
  const int b = blockIdx.z;
  const bool align = flags & OCF_WARP_ALIGN_CORNERS;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;

  int pix[R];
  bool act[R];
  float wnw[R], wne[R], wsw[R], wse[R], wx0[R], wx1[R], wy0[R], wy1[R], gmul[R], mul[R];
  int onw[R], one[R], osw[R], ose[R];  // -1 when the tap is dropped
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int y = yg * R + r;
    act[r] = in_grid && y < H;
    const int yc = min(y, H - 1);
    pix[r] = yc * W + x;
    const float u = __fmul_rn(__ldg(flow + ((size_t)b * 2) * HW + pix[r]), scale);
    const float v = __fmul_rn(__ldg(flow + ((size_t)b * 2 + 1) * HW + pix[r]), scale);
    const float ix = unnormalize(__fadd_rn((float)x, u), W, max(W - 1, 1), align);
    const float iy = unnormalize(__fadd_rn((float)yc, v), H, max(H - 1, 1), align);
    const Taps t = make_taps(ix, iy, H, W);
    mul[r] = act[r] ? 1.f : 0.f;
    if (flags & OCF_WARP_IS_MASK) mul[r] *= mask_of(t);
    const float occv = occ != nullptr ? __ldg(occ + (size_t)b * HW + pix[r]) : 1.f;
    gmul[r] = mul[r] * occv;  // d out / d sample
    wx0[r] = t.wx0; wx1[r] = t.wx1; wy0[r] = t.wy0; wy1[r] = t.wy1;
    wnw[r] = t.wx0 * t.wy0; wne[r] = t.wx1 * t.wy0; wsw[r] = t.wx0 * t.wy1; wse[r] = t.wx1 * t.wy1;
    const int o = t.y0 * W + t.x0;
    onw[r] = (act[r] && t.vx0 && t.vy0) ? o : -1;
    one[r] = (act[r] && t.vx1 && t.vy0) ? o + 1 : -1;
    osw[r] = (act[r] && t.vx0 && t.vy1) ? o + W : -1;
    ose[r] = (act[r] && t.vx1 && t.vy1) ? o + W + 1 : -1;
  }
  // merge plan (channel independent)
  bool mvw[R], mve[R], emit_sw[R], emit_se[R], take_n[R], give_n[R], take_s[R], give_s[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    mvw[r] = r > 0 && onw[r] >= 0 && onw[r] == osw[r - 1];   // north-west tap of row r absorbs south-west of row r-1
    mve[r] = r > 0 && one[r] >= 0 && one[r] == ose[r - 1];
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    emit_sw[r] = osw[r] >= 0 && !(r + 1 < R && mvw[r + 1]);
    emit_se[r] = ose[r] >= 0 && !(r + 1 < R && mve[r + 1]);
    const int pne = __shfl_up_sync(full, one[r], 1);
    const int pse = __shfl_up_sync(full, emit_se[r] ? ose[r] : -1, 1);
    take_n[r] = lane > 0 && onw[r] >= 0 && pne == onw[r];
    take_s[r] = lane > 0 && emit_sw[r] && pse == osw[r];
    give_n[r] = __shfl_down_sync(full, (int)take_n[r], 1) && lane < 31;
    give_s[r] = __shfl_down_sync(full, (int)take_s[r], 1) && lane < 31;
  }

  const int c_begin = blockIdx.y * slab, c_end = min(C, c_begin + slab);
  const float* ip = img + ((size_t)b * C + c_begin) * HW;
  const float* gp = gout + ((size_t)b * C + c_begin) * HW;
  float* dp = d_img != nullptr ? d_img + ((size_t)b * C + c_begin) * HW : nullptr;
  float gx[R], gy[R], go[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { gx[r] = 0.f; gy[r] = 0.f; go[r] = 0.f; }
  const bool need_vals = d_flow != nullptr || d_occ != nullptr;
  // Channels go in batches of CB: all loads of a batch are issued before its first atomic (atomics are ordering points
  // for the compiler, so without the explicit batch every channel would expose a full global-load latency).
  constexpr int CB = OCF_WB_CB;
  for (int c0 = c_begin; c0 < c_end; c0 += CB, ip += (size_t)CB * HW, gp += (size_t)CB * HW) {
    float gv[CB][R], ta[CB][R], tb[CB][R], tc[CB][R], td[CB][R];
#pragma unroll
    for (int j = 0; j < CB; ++j) {
      const bool cok = c0 + j < c_end;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        gv[j][r] = (cok && act[r]) ? __ldg(gp + (size_t)j * HW + pix[r]) : 0.f;
        if (need_vals) {
          ta[j][r] = (cok && onw[r] >= 0) ? __ldg(ip + (size_t)j * HW + onw[r]) : 0.f;
          tb[j][r] = (cok && one[r] >= 0) ? __ldg(ip + (size_t)j * HW + one[r]) : 0.f;
          tc[j][r] = (cok && osw[r] >= 0) ? __ldg(ip + (size_t)j * HW + osw[r]) : 0.f;
          td[j][r] = (cok && ose[r] >= 0) ? __ldg(ip + (size_t)j * HW + ose[r]) : 0.f;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < CB; ++j) {
      if (c0 + j >= c_end) break;  // uniform
      float cnw[R], cne[R], csw[R], cse[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float graw = gv[j][r];
        const float g = graw * gmul[r];
        if (need_vals) {
          const float a = ta[j][r], bb = tb[j][r], cc = tc[j][r], dd = td[j][r];
          // ATen grid_sampler_2d_backward: gix -= nw*(iy_se-iy) ; += ne*(iy_sw-iy) ; -= sw*(iy-iy_ne) ; += se*(iy-iy_nw)
          gx[r] += g * ((bb - a) * wy0[r] + (dd - cc) * wy1[r]);
          gy[r] += g * ((cc - a) * wx0[r] + (dd - bb) * wx1[r]);
          if (d_occ != nullptr) go[r] += graw * mul[r] * (a * wnw[r] + bb * wne[r] + cc * wsw[r] + dd * wse[r]);
        }
        cnw[r] = g * wnw[r]; cne[r] = g * wne[r]; csw[r] = g * wsw[r]; cse[r] = g * wse[r];
      }
      if (dp != nullptr) {
#pragma unroll
        for (int r = 1; r < R; ++r) {
          if (mvw[r]) cnw[r] += csw[r - 1];
          if (mve[r]) cne[r] += cse[r - 1];
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float pn = __shfl_up_sync(full, cne[r], 1), ps = __shfl_up_sync(full, cse[r], 1);
          if (take_n[r]) cnw[r] += pn;
          if (take_s[r]) csw[r] += ps;
          if (onw[r] >= 0) atomicAdd(dp + onw[r], cnw[r]);
          if (one[r] >= 0 && !give_n[r]) atomicAdd(dp + one[r], cne[r]);
          if (emit_sw[r]) atomicAdd(dp + osw[r], csw[r]);
          if (emit_se[r] && !give_s[r]) atomicAdd(dp + ose[r], cse[r]);
        }
        dp += HW;
      }
    }
  }
  // chain: ATen multiplies by (W-1)/2 resp. W/2, the reference's normalisation by 2/max(W-1,1)
  const float mx = (align ? 0.5f * (float)(W - 1) : 0.5f * (float)W) * (2.0f / (float)max(W - 1, 1)) * scale;
  const float my = (align ? 0.5f * (float)(H - 1) : 0.5f * (float)H) * (2.0f / (float)max(H - 1, 1)) * scale;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (!act[r]) continue;
    if (d_flow != nullptr) {
      float* fx = d_flow + ((size_t)b * 2) * HW + pix[r];
      if (nslabs == 1) { fx[0] = gx[r] * mx; fx[HW] = gy[r] * my; }
      else { atomicAdd(fx, gx[r] * mx); atomicAdd(fx + HW, gy[r] * my); }
    }
    if (d_occ != nullptr) {
      float* po = d_occ + (size_t)b * HW + pix[r];
      if (nslabs == 1) *po = go[r]; else atomicAdd(po, go[r]);
    }
  }
}

// ---- quad path (W % 4 == 0) ------------------------------------------------------------------------
// A thread owns 4 horizontally adjacent samples.  For a spatially coherent flow (what the decoders produce: an
// up-sampled coarse field) their north-west taps are 4 consecutive pixels of ONE source row, so the 5 + 5 source
// pixels of the two tap rows sit inside two 16-byte aligned groups per row: the gather is 4 x LDG.128 and the
// backward scatter 4 x RED.128 (red.global.add.v4.f32) per 4 samples and channel instead of 16 scalar loads /
// 16 scalar reds.  The lane-dependent position inside the aligned groups (o = x0 & 3) is resolved with a 2-stage
// register funnel (select chains, no local memory).  Quads that are not coherent (flow discontinuities, samples
// leaving the frame through different rows) take the scalar per-tap path inside the same kernel.
// The L2 atomic unit is fed per lane-operation (~1.3 cycles per lane and SM for scalar reds, measured in round 1:
// warp_bwd and range_map ran exactly at that rate), so 4x fewer red operations is 4x less time in the scatter.
constexpr int QUAD_THREADS = 256;

struct QuadPlan {
  bool fast;        // coherent: one tap row pair, consecutive columns
  bool nA, nB, sA, sB;  // aligned group A = [a, a+3] / B = [a+4, a+7] of the north / south tap row inside the image
  int o;            // x0[0] - a
  int rowN;         // plane offset of group A in the north tap row (south: + W)
};

__device__ __forceinline__ QuadPlan make_plan(const Taps (&t)[4], int H, int W) {
  QuadPlan p;
  p.fast = t[0].x0 >= -8 && t[0].x0 <= W && t[0].y0 >= -2 && t[0].y0 <= H;   // keeps the offset arithmetic in range
#pragma unroll
  for (int j = 1; j < 4; ++j) p.fast = p.fast && t[j].y0 == t[0].y0 && t[j].x0 == t[0].x0 + j;
  const int a = t[0].x0 & ~3;  // floor to a multiple of 4 (two's complement: also for negative x0)
  p.o = t[0].x0 - a;
  const bool vn = t[0].y0 >= 0 && t[0].y0 < H, vs = t[0].y0 + 1 >= 0 && t[0].y0 + 1 < H;
  const bool va = a >= 0 && a + 3 < W, vb = a + 4 >= 0 && a + 7 < W;   // W % 4 == 0: a group is entirely inside or outside
  p.nA = vn && va; p.nB = vn && vb; p.sA = vs && va; p.sB = vs && vb;
  p.rowN = t[0].y0 * W + a;
  return p;
}

__device__ __forceinline__ void quad_taps(const float* __restrict__ flow, int b, int HW, int pix, int x, int y, int H, int W,
                                          bool align, float scale, Taps (&t)[4]) {
  const float4 u4 = __ldg(reinterpret_cast<const float4*>(flow + ((size_t)b * 2) * HW + pix));
  const float4 v4 = __ldg(reinterpret_cast<const float4*>(flow + ((size_t)b * 2 + 1) * HW + pix));
  const float us[4] = {u4.x, u4.y, u4.z, u4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float ix = unnormalize(__fadd_rn((float)(x + j), __fmul_rn(us[j], scale)), W, max(W - 1, 1), align);
    const float iy = unnormalize(__fadd_rn((float)y, __fmul_rn(vs[j], scale)), H, max(H - 1, 1), align);
    t[j] = make_taps(ix, iy, H, W);
  }
}

// grid: (ceil(H*W/4 / 256), channel slabs, B)
__global__ void __launch_bounds__(QUAD_THREADS)
warp_fwd_quad(const float* __restrict__ img, const float* __restrict__ flow, const float* __restrict__ occ,
              float* __restrict__ out, int C, int H, int W, int slab, int flags, float scale) {
  const int W4 = W >> 2, HW = H * W;
  const int q = blockIdx.x * QUAD_THREADS + threadIdx.x;
  if (q >= H * W4) return;
  const int y = q / W4, x = (q - y * W4) << 2;
  const int b = blockIdx.z, pix = y * W + x;
  Taps t[4];
  quad_taps(flow, b, HW, pix, x, y, H, W, flags & OCF_WARP_ALIGN_CORNERS, scale, t);
  float mul[4] = {1.f, 1.f, 1.f, 1.f};
  if (flags & OCF_WARP_IS_MASK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) mul[j] = mask_of(t[j]);
  }
  if (occ != nullptr) {
    const float4 o4 = __ldg(reinterpret_cast<const float4*>(occ + (size_t)b * HW + pix));
    mul[0] *= o4.x; mul[1] *= o4.y; mul[2] *= o4.z; mul[3] *= o4.w;
  }
  const QuadPlan pl = make_plan(t, H, W);
  const int c_begin = blockIdx.y * slab, c_end = min(C, c_begin + slab);
  const float* ip = img + ((size_t)b * C + c_begin) * HW;
  float* op = out + ((size_t)b * C + c_begin) * HW + pix;
  if (pl.fast) {
    float wnw[4], wne[4], wsw[4], wse[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      wnw[j] = t[j].wx0 * t[j].wy0; wne[j] = t[j].wx1 * t[j].wy0; wsw[j] = t[j].wx0 * t[j].wy1; wse[j] = t[j].wx1 * t[j].wy1;
    }
    const float* rp = ip + pl.rowN;
#pragma unroll 2
    for (int c = c_begin; c < c_end; ++c, rp += HW, op += HW) {
      const float4 a0 = ldg4_or_zero(rp, pl.nA), a1 = ldg4_or_zero(rp + 4, pl.nB);
      const float4 b0 = ldg4_or_zero(rp + W, pl.sA), b1 = ldg4_or_zero(rp + W + 4, pl.sB);
      const float qn[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float qs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float vn[5], vs[5], r[4];
      funnel_gather(qn, pl.o, vn);
      funnel_gather(qs, pl.o, vs);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float acc = 0.f;   // same tap order as the scalar kernel: nw, ne, sw, se
        acc = fmaf(vn[j], wnw[j], acc);
        acc = fmaf(vn[j + 1], wne[j], acc);
        acc = fmaf(vs[j], wsw[j], acc);
        acc = fmaf(vs[j + 1], wse[j], acc);
        r[j] = acc * mul[j];
      }
      *reinterpret_cast<float4*>(op) = make_float4(r[0], r[1], r[2], r[3]);
    }
  } else {
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c, ip += HW, op += HW) {
      float r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const Taps& tj = t[j];
        const int o = tj.y0 * W + tj.x0;
        float acc = 0.f;
        if (tj.vx0 && tj.vy0) acc = fmaf(__ldg(ip + o), tj.wx0 * tj.wy0, acc);
        if (tj.vx1 && tj.vy0) acc = fmaf(__ldg(ip + o + 1), tj.wx1 * tj.wy0, acc);
        if (tj.vx0 && tj.vy1) acc = fmaf(__ldg(ip + o + W), tj.wx0 * tj.wy1, acc);
        if (tj.vx1 && tj.vy1) acc = fmaf(__ldg(ip + o + W + 1), tj.wx1 * tj.wy1, acc);
        r[j] = acc * mul[j];
      }
      *reinterpret_cast<float4*>(op) = make_float4(r[0], r[1], r[2], r[3]);
    }
  }
}

// Backward, quad path.  d_img is zeroed by the entry point; d_flow / d_occ are per-sample and accumulated across channel
// slabs with atomics (plain 128-bit stores when there is one slab).
__global__ void __launch_bounds__(QUAD_THREADS, 2)
warp_bwd_quad(const float* __restrict__ gout, const float* __restrict__ img, const float* __restrict__ flow,
              const float* __restrict__ occ, float* __restrict__ d_img, float* __restrict__ d_flow,
              float* __restrict__ d_occ, int C, int H, int W, int slab, int nslabs, int flags, float scale) {
  const int W4 = W >> 2, HW = H * W;
  const int q = blockIdx.x * QUAD_THREADS + threadIdx.x;
  if (q >= H * W4) return;
  const int y = q / W4, x = (q - y * W4) << 2;
  const int b = blockIdx.z, pix = y * W + x;
  const bool align = flags & OCF_WARP_ALIGN_CORNERS;
  Taps t[4];
  quad_taps(flow, b, HW, pix, x, y, H, W, align, scale, t);
  float mul[4] = {1.f, 1.f, 1.f, 1.f}, gmul[4];
  if (flags & OCF_WARP_IS_MASK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) mul[j] = mask_of(t[j]);
  }
  gmul[0] = mul[0]; gmul[1] = mul[1]; gmul[2] = mul[2]; gmul[3] = mul[3];
  if (occ != nullptr) {
    const float4 o4 = __ldg(reinterpret_cast<const float4*>(occ + (size_t)b * HW + pix));
    gmul[0] *= o4.x; gmul[1] *= o4.y; gmul[2] *= o4.z; gmul[3] *= o4.w;
  }
  const QuadPlan pl = make_plan(t, H, W);
  const int c_begin = blockIdx.y * slab, c_end = min(C, c_begin + slab);
  const float* ip = img + ((size_t)b * C + c_begin) * HW;
  const float* gp = gout + ((size_t)b * C + c_begin) * HW + pix;
  float* dp = d_img != nullptr ? d_img + ((size_t)b * C + c_begin) * HW : nullptr;
  const bool need_vals = d_flow != nullptr || d_occ != nullptr;
  float gx[4] = {0.f, 0.f, 0.f, 0.f}, gy[4] = {0.f, 0.f, 0.f, 0.f}, go[4] = {0.f, 0.f, 0.f, 0.f};
  float wnw[4], wne[4], wsw[4], wse[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    wnw[j] = t[j].wx0 * t[j].wy0; wne[j] = t[j].wx1 * t[j].wy0; wsw[j] = t[j].wx0 * t[j].wy1; wse[j] = t[j].wx1 * t[j].wy1;
  }
  if (pl.fast) {
    const float* rp = ip + pl.rowN;
    float* drp = dp != nullptr ? dp + pl.rowN : nullptr;
#pragma unroll 2
    for (int c = c_begin; c < c_end; ++c, rp += HW, gp += HW) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gp));
      const float graw[4] = {g4.x, g4.y, g4.z, g4.w};
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = graw[j] * gmul[j];
      if (need_vals) {
        const float4 a0 = ldg4_or_zero(rp, pl.nA), a1 = ldg4_or_zero(rp + 4, pl.nB);
        const float4 b0 = ldg4_or_zero(rp + W, pl.sA), b1 = ldg4_or_zero(rp + W + 4, pl.sB);
        const float qn[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float qs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float vn[5], vs[5];
        funnel_gather(qn, pl.o, vn);
        funnel_gather(qs, pl.o, vs);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = vn[j], bb = vn[j + 1], cc = vs[j], dd = vs[j + 1];
          gx[j] += g[j] * ((bb - a) * t[j].wy0 + (dd - cc) * t[j].wy1);
          gy[j] += g[j] * ((cc - a) * t[j].wx0 + (dd - bb) * t[j].wx1);
          if (d_occ != nullptr) go[j] += graw[j] * mul[j] * (a * wnw[j] + bb * wne[j] + cc * wsw[j] + dd * wse[j]);
        }
      }
      if (drp != nullptr) {
        float cn[5], cs[5], qn[8], qs[8];
        cn[0] = g[0] * wnw[0]; cs[0] = g[0] * wsw[0];
#pragma unroll
        for (int j = 1; j < 4; ++j) {
          cn[j] = fmaf(g[j], wnw[j], g[j - 1] * wne[j - 1]);
          cs[j] = fmaf(g[j], wsw[j], g[j - 1] * wse[j - 1]);
        }
        cn[4] = g[3] * wne[3]; cs[4] = g[3] * wse[3];
        funnel_scatter(cn, pl.o, qn);
        funnel_scatter(cs, pl.o, qs);
        if (pl.nA) red_add_v4(drp, qn[0], qn[1], qn[2], qn[3]);
        if (pl.nB) red_add_v4(drp + 4, qn[4], qn[5], qn[6], qn[7]);
        if (pl.sA) red_add_v4(drp + W, qs[0], qs[1], qs[2], qs[3]);
        if (pl.sB) red_add_v4(drp + W + 4, qs[4], qs[5], qs[6], qs[7]);
        drp += HW;
      }
    }
  } else {
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c, ip += HW, gp += HW) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gp));
      const float graw[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const Taps& tj = t[j];
        const int o = tj.y0 * W + tj.x0;
        const bool vnw = tj.vx0 && tj.vy0, vne = tj.vx1 && tj.vy0, vsw = tj.vx0 && tj.vy1, vse = tj.vx1 && tj.vy1;
        const float g = graw[j] * gmul[j];
        if (need_vals) {
          const float a = vnw ? __ldg(ip + o) : 0.f, bb = vne ? __ldg(ip + o + 1) : 0.f;
          const float cc = vsw ? __ldg(ip + o + W) : 0.f, dd = vse ? __ldg(ip + o + W + 1) : 0.f;
          gx[j] += g * ((bb - a) * tj.wy0 + (dd - cc) * tj.wy1);
          gy[j] += g * ((cc - a) * tj.wx0 + (dd - bb) * tj.wx1);
          if (d_occ != nullptr) go[j] += graw[j] * mul[j] * (a * wnw[j] + bb * wne[j] + cc * wsw[j] + dd * wse[j]);
        }
        if (dp != nullptr) {
          if (vnw) atomicAdd(dp + o, g * wnw[j]);
          if (vne) atomicAdd(dp + o + 1, g * wne[j]);
          if (vsw) atomicAdd(dp + o + W, g * wsw[j]);
          if (vse) atomicAdd(dp + o + W + 1, g * wse[j]);
        }
      }
      if (dp != nullptr) dp += HW;
    }
  }
  const float mx = (align ? 0.5f * (float)(W - 1) : 0.5f * (float)W) * (2.0f / (float)max(W - 1, 1)) * scale;
  const float my = (align ? 0.5f * (float)(H - 1) : 0.5f * (float)H) * (2.0f / (float)max(H - 1, 1)) * scale;
  if (d_flow != nullptr) {
    float* fx = d_flow + ((size_t)b * 2) * HW + pix;
    if (nslabs == 1) {
      *reinterpret_cast<float4*>(fx) = make_float4(gx[0] * mx, gx[1] * mx, gx[2] * mx, gx[3] * mx);
      *reinterpret_cast<float4*>(fx + HW) = make_float4(gy[0] * my, gy[1] * my, gy[2] * my, gy[3] * my);
    } else {
      red_add_v4(fx, gx[0] * mx, gx[1] * mx, gx[2] * mx, gx[3] * mx);
      red_add_v4(fx + HW, gy[0] * my, gy[1] * my, gy[2] * my, gy[3] * my);
    }
  }
  if (d_occ != nullptr) {
    float* po = d_occ + (size_t)b * HW + pix;
    if (nslabs == 1) *reinterpret_cast<float4*>(po) = make_float4(go[0], go[1], go[2], go[3]);
    else red_add_v4(po, go[0], go[1], go[2], go[3]);
  }
}

// Range map, quad path: the 2 x 5 bilinear splat weights of 4 coherent source pixels as 4 x RED.128.
__global__ void __launch_bounds__(QUAD_THREADS)
range_map_quad(const float* __restrict__ flow, float* __restrict__ range, int H, int W) {
  const int W4 = W >> 2, HW = H * W;
  const int q = blockIdx.x * QUAD_THREADS + threadIdx.x;
  if (q >= H * W4) return;
  const int y = q / W4, x = (q - y * W4) << 2;
  const int b = blockIdx.y, pix = y * W + x;
  const float4 u4 = __ldg(reinterpret_cast<const float4*>(flow + ((size_t)b * 2) * HW + pix));
  const float4 v4 = __ldg(reinterpret_cast<const float4*>(flow + ((size_t)b * 2 + 1) * HW + pix));
  const float us[4] = {u4.x, u4.y, u4.z, u4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w};
  int x0[4], y0[4];
  float w00[4], w10[4], w01[4], w11[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float ex = __fadd_rn((float)(x + j), us[j]);   // flow_to_warp, model.py:223-241
    const float ey = __fadd_rn((float)y, vs[j]);
    const float fx0 = floorf(ex), fy0 = floorf(ey);
    const float ox = ex - fx0, oy = ey - fy0;
    const bool sane = fabsf(fx0) < 1.0e9f && fabsf(fy0) < 1.0e9f;   // the reference's .to(int32) is undefined for huge values
    x0[j] = sane ? (int)fx0 : -10;
    y0[j] = sane ? (int)fy0 : -10;
    w00[j] = (1.f - ox) * (1.f - oy); w10[j] = ox * (1.f - oy); w01[j] = (1.f - ox) * oy; w11[j] = ox * oy;
  }
  float* r = range + (size_t)b * HW;
  bool fast = x0[0] >= -8 && x0[0] <= W && y0[0] >= -2 && y0[0] <= H;
#pragma unroll
  for (int j = 1; j < 4; ++j) fast = fast && y0[j] == y0[0] && x0[j] == x0[0] + j;
  if (fast) {
    const int a = x0[0] & ~3, o = x0[0] - a;
    const bool vn = y0[0] >= 0 && y0[0] < H, vs2 = y0[0] + 1 >= 0 && y0[0] + 1 < H;
    const bool va = a >= 0 && a + 3 < W, vb = a + 4 >= 0 && a + 7 < W;
    float cn[5], cs[5], qn[8], qs[8];
    cn[0] = w00[0]; cs[0] = w01[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) { cn[j] = w00[j] + w10[j - 1]; cs[j] = w01[j] + w11[j - 1]; }
    cn[4] = w10[3]; cs[4] = w11[3];
    funnel_scatter(cn, o, qn);
    funnel_scatter(cs, o, qs);
    float* rp = r + y0[0] * W + a;
    if (vn && va) red_add_v4(rp, qn[0], qn[1], qn[2], qn[3]);
    if (vn && vb) red_add_v4(rp + 4, qn[4], qn[5], qn[6], qn[7]);
    if (vs2 && va) red_add_v4(rp + W, qs[0], qs[1], qs[2], qs[3]);
    if (vs2 && vb) red_add_v4(rp + W + 4, qs[4], qs[5], qs[6], qs[7]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool vx0 = x0[j] >= 0 && x0[j] < W, vx1 = x0[j] + 1 >= 0 && x0[j] + 1 < W;
      const bool vy0 = y0[j] >= 0 && y0[j] < H, vy1 = y0[j] + 1 >= 0 && y0[j] + 1 < H;
      const int o = y0[j] * W + x0[j];
      if (vx0 && vy0) atomicAdd(r + o, w00[j]);
      if (vx1 && vy0) atomicAdd(r + o + 1, w10[j]);
      if (vx0 && vy1) atomicAdd(r + o + W, w01[j]);
      if (vx1 && vy1) atomicAdd(r + o + W + 1, w11[j]);
    }
  }
}

// ---- range map ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
range_map_kernel(const float* __restrict__ flow, float* __restrict__ range, int H, int W) {
  const int pix = blockIdx.x * 256 + threadIdx.x;
  const int HW = H * W;
  const bool active = pix < HW;
  const int pixc = active ? pix : HW - 1;
  const int y = pixc / W, x = pixc - y * W;
  const int b = blockIdx.y;
  const float ex = __fadd_rn((float)x, __ldg(flow + ((size_t)b * 2) * HW + pixc));      // flow_to_warp, model.py:223-241
  const float ey = __fadd_rn((float)y, __ldg(flow + ((size_t)b * 2 + 1) * HW + pixc));
  const float fx0 = floorf(ex), fy0 = floorf(ey);
  const float ox = ex - fx0, oy = ey - fy0;
  // guard the float->int conversion (the reference's .to(int32) is undefined for huge values)
  const bool sane = active && fabsf(fx0) < 1.0e9f && fabsf(fy0) < 1.0e9f;
  const int x0 = sane ? (int)fx0 : -10, y0 = sane ? (int)fy0 : -10;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
  const bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
  float w00 = (1.f - ox) * (1.f - oy), w10 = ox * (1.f - oy), w01 = (1.f - ox) * oy, w11 = ox * oy;
  const int o00 = (vx0 && vy0) ? y0 * W + x0 : -1, o10 = (vx1 && vy0) ? y0 * W + x0 + 1 : -1;
  const int o01 = (vx0 && vy1) ? (y0 + 1) * W + x0 : -1, o11 = (vx1 && vy1) ? (y0 + 1) * W + x0 + 1 : -1;
  // warp-aggregated scatter: fold the previous lane's east taps into my west taps when they coincide
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int p10 = __shfl_up_sync(full, o10, 1), p11 = __shfl_up_sync(full, o11, 1);
  const float q10 = __shfl_up_sync(full, w10, 1), q11 = __shfl_up_sync(full, w11, 1);
  const bool take0 = lane > 0 && o00 >= 0 && p10 == o00;
  const bool take1 = lane > 0 && o01 >= 0 && p11 == o01;
  if (take0) w00 += q10;
  if (take1) w01 += q11;
  const bool give0 = __shfl_down_sync(full, (int)take0, 1) && lane < 31;
  const bool give1 = __shfl_down_sync(full, (int)take1, 1) && lane < 31;
  float* r = range + (size_t)b * HW;
  if (o00 >= 0) atomicAdd(r + o00, w00);
  if (o10 >= 0 && !give0) atomicAdd(r + o10, w10);
  if (o01 >= 0) atomicAdd(r + o01, w01);
  if (o11 >= 0 && !give1) atomicAdd(r + o11, w11);
}

__global__ void __launch_bounds__(256)
occ_from_range_kernel(const float* __restrict__ range, float* __restrict__ occ, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) occ[i] = 1.0f - fminf(fmaxf(range[i], 0.0f), 1.0f);  // model.py:391
}

__global__ void __launch_bounds__(256)
flow_to_warp_kernel(const float* __restrict__ flow, float* __restrict__ out, int H, int W, size_t n2) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;  // index over B*H*W*2 (channels-last flow)
  if (i >= n2) return;
  const size_t p = i >> 1;
  const int comp = (int)(i & 1);
  const int x = (int)(p % W), y = (int)((p / W) % H);
  out[i] = __fadd_rn(comp == 0 ? (float)x : (float)y, flow[i]);
}

int pick_slab(int C, int HW, int B) {
  // enough CTAs to fill 148 SMs a few times over, but at most 16 channels of coordinate reuse per thread
  const long long pix_blocks = ((long long)HW + WARP_THREADS - 1) / WARP_THREADS * B;
  int slab = C;
  while (slab > 4 && pix_blocks * ((C + slab - 1) / slab) < 4LL * OCF_SM_COUNT) slab = (slab + 1) / 2;
  if (slab > 32) slab = 32;
  return slab < 1 ? 1 : slab;
}

// developer knob for tuning runs: OCF_WARP_QUAD bit 0 = forward, bit 1 = backward, bit 2 = range map ; OCF_QUAD_CTAS = CTAs per SM aimed at
int quad_mask() {
  static const int m = []() { const char* e = getenv("OCF_WARP_QUAD"); return e ? atoi(e) : 0; }();
  return m;
}
int quad_ctas() {
  static const int m = []() { const char* e = getenv("OCF_QUAD_CTAS"); return e ? atoi(e) : 2; }();
  return m;
}

// quad kernels: a thread covers 4 samples, so 4x fewer threads per channel slab; aim for >= 4 CTAs per SM
int pick_slab_quad(int C, int HW, int B) {
  const long long blocks = ((long long)(HW / 4) + QUAD_THREADS - 1) / QUAD_THREADS * B;
  int slab = C;
  while (slab > 4 && blocks * ((C + slab - 1) / slab) < (long long)quad_ctas() * OCF_SM_COUNT) slab = (slab + 1) / 2;
  if (slab > 32) slab = 32;
  return slab < 1 ? 1 : slab;
}

bool quad_ok(int W, const void* a, const void* b, const void* c, const void* d, const void* e, const void* f, const void* g) {
  const void* ps[7] = {a, b, c, d, e, f, g};
  if (W % 4 != 0) return false;
  for (const void* p : ps)
    if (p != nullptr && !ocf_aligned16(p)) return false;
  return true;
}

}  // namespace

extern "C" int ocf_warp_fwd(const float* img, const float* flow, const float* occ, float* out, int B, int C, int H, int W,
                            int flags, float scale, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(img); OCF_REQUIRE_PTR(flow); OCF_REQUIRE_PTR(out);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE((long long)H * W < (1LL << 30) && B <= 65535, OCF_EUNSUPPORTED);
  OCF_REQUIRE((flags & ~3) == 0, OCF_EUNSUPPORTED);
  const int HW = H * W;
  if ((quad_mask() & 1) && quad_ok(W, img, flow, occ, out, nullptr, nullptr, nullptr)) {
    const int slab = pick_slab_quad(C, HW, B);
    const int nslabs = (C + slab - 1) / slab;
    OCF_REQUIRE(nslabs <= 65535, OCF_EUNSUPPORTED);
    dim3 grid((HW / 4 + QUAD_THREADS - 1) / QUAD_THREADS, nslabs, B);
    warp_fwd_quad<<<grid, QUAD_THREADS, 0, ocf_cast_stream(stream)>>>(img, flow, occ, out, C, H, W, slab, flags, scale);
    return ocf_launch_status();
  }
  const int slab = pick_slab(C, HW, B);
  const int nslabs = (C + slab - 1) / slab;
  OCF_REQUIRE(nslabs <= 65535, OCF_EUNSUPPORTED);
  dim3 grid((HW + WARP_THREADS - 1) / WARP_THREADS, nslabs, B);
  warp_fwd_kernel<<<grid, WARP_THREADS, 0, ocf_cast_stream(stream)>>>(img, flow, occ, out, C, H, W, slab, flags, scale);
  return ocf_launch_status();
}

extern "C" int ocf_warp_bwd(const float* grad_out, const float* img, const float* flow, const float* occ, float* d_img,
                            float* d_flow, float* d_occ, int B, int C, int H, int W, int flags, float scale,
                            ocf_stream_t stream) {
  OCF_REQUIRE_PTR(grad_out); OCF_REQUIRE_PTR(img); OCF_REQUIRE_PTR(flow);
  OCF_REQUIRE(d_img != nullptr || d_flow != nullptr || d_occ != nullptr, OCF_ENULL);
  OCF_REQUIRE(d_occ == nullptr || occ != nullptr, OCF_ENULL);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE((long long)H * W < (1LL << 30) && B <= 65535, OCF_EUNSUPPORTED);
  OCF_REQUIRE((flags & ~3) == 0, OCF_EUNSUPPORTED);
  cudaStream_t s = ocf_cast_stream(stream);
  const int HW = H * W;
  const bool quad = (quad_mask() & 2) && quad_ok(W, grad_out, img, flow, occ, d_img, d_flow, d_occ);
  const int slab = quad ? pick_slab_quad(C, HW, B) : pick_slab(C, HW, B);
  const int nslabs = (C + slab - 1) / slab;
  OCF_REQUIRE(nslabs <= 65535, OCF_EUNSUPPORTED);
  cudaError_t e;
  if (d_img != nullptr && (e = cudaMemsetAsync(d_img, 0, sizeof(float) * (size_t)B * C * HW, s)) != cudaSuccess) return (int)e;
  if (nslabs > 1) {
    if (d_flow != nullptr && (e = cudaMemsetAsync(d_flow, 0, sizeof(float) * (size_t)B * 2 * HW, s)) != cudaSuccess) return (int)e;
    if (d_occ != nullptr && (e = cudaMemsetAsync(d_occ, 0, sizeof(float) * (size_t)B * HW, s)) != cudaSuccess) return (int)e;
  }
  if (quad) {
    dim3 qgrid((HW / 4 + QUAD_THREADS - 1) / QUAD_THREADS, nslabs, B);
    warp_bwd_quad<<<qgrid, QUAD_THREADS, 0, s>>>(grad_out, img, flow, occ, d_img, d_flow, d_occ, C, H, W, slab, nslabs, flags, scale);
    return ocf_launch_status();
  }
  constexpr int R = OCF_WB_R;
  dim3 grid(((H + R - 1) / R * W + WARP_THREADS - 1) / WARP_THREADS, nslabs, B);
  warp_bwd_kernel<R><<<grid, WARP_THREADS, 0, s>>>(grad_out, img, flow, occ, d_img, d_flow, d_occ, C, H, W, slab, nslabs, flags, scale);
  return ocf_launch_status();
}

extern "C" int ocf_range_map(const float* flow, float* range_out, float* occ_out, int B, int H, int W, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(flow); OCF_REQUIRE_PTR(range_out);
  OCF_REQUIRE(B > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE((long long)H * W < (1LL << 30) && B <= 65535, OCF_EUNSUPPORTED);
  cudaStream_t s = ocf_cast_stream(stream);
  const int HW = H * W;
  cudaError_t e = cudaMemsetAsync(range_out, 0, sizeof(float) * (size_t)B * HW, s);
  if (e != cudaSuccess) return (int)e;
  if ((quad_mask() & 4) && quad_ok(W, flow, range_out, nullptr, nullptr, nullptr, nullptr, nullptr)) {
    dim3 qgrid((HW / 4 + QUAD_THREADS - 1) / QUAD_THREADS, B);
    range_map_quad<<<qgrid, QUAD_THREADS, 0, s>>>(flow, range_out, H, W);
  } else {
    dim3 grid((HW + 255) / 256, B);
    range_map_kernel<<<grid, 256, 0, s>>>(flow, range_out, H, W);
  }
  if (int st = ocf_launch_status()) return st;
  if (occ_out != nullptr) {
    const size_t n = (size_t)B * HW;
    occ_from_range_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(range_out, occ_out, n);
  }
  return ocf_launch_status();
}

extern "C" int ocf_flow_to_warp(const float* flow_bhw2, float* out_bhw2, int B, int H, int W, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(flow_bhw2); OCF_REQUIRE_PTR(out_bhw2);
  OCF_REQUIRE(B > 0 && H > 0 && W > 0, OCF_ESHAPE);
  const size_t n2 = (size_t)B * H * W * 2;
  OCF_REQUIRE(n2 < (1ULL << 39), OCF_EUNSUPPORTED);
  flow_to_warp_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, ocf_cast_stream(stream)>>>(flow_bhw2, out_bhw2, H, W, n2);
  return ocf_launch_status();
}
