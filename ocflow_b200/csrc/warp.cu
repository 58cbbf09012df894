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

}  // namespace

extern "C" int ocf_warp_fwd(const float* img, const float* flow, const float* occ, float* out, int B, int C, int H, int W,
                            int flags, float scale, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(img); OCF_REQUIRE_PTR(flow); OCF_REQUIRE_PTR(out);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE((long long)H * W < (1LL << 30) && B <= 65535, OCF_EUNSUPPORTED);
  OCF_REQUIRE((flags & ~3) == 0, OCF_EUNSUPPORTED);
  const int HW = H * W;
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
  const int slab = pick_slab(C, HW, B);
  const int nslabs = (C + slab - 1) / slab;
  OCF_REQUIRE(nslabs <= 65535, OCF_EUNSUPPORTED);
  cudaError_t e;
  if (d_img != nullptr && (e = cudaMemsetAsync(d_img, 0, sizeof(float) * (size_t)B * C * HW, s)) != cudaSuccess) return (int)e;
  if (nslabs > 1) {
    if (d_flow != nullptr && (e = cudaMemsetAsync(d_flow, 0, sizeof(float) * (size_t)B * 2 * HW, s)) != cudaSuccess) return (int)e;
    if (d_occ != nullptr && (e = cudaMemsetAsync(d_occ, 0, sizeof(float) * (size_t)B * HW, s)) != cudaSuccess) return (int)e;
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
  dim3 grid((HW + 255) / 256, B);
  range_map_kernel<<<grid, 256, 0, s>>>(flow, range_out, H, W);
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
