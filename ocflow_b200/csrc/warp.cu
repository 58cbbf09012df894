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

// Both warp kernels are latency-bound unless many loads are in flight per thread (ncu, profiles/r2_stream_*: long_scoreboard
// dominated at 1-2 channels per batch), so the channel loop runs in batches of CB channels whose 4 x CB tap loads (and CB
// gradient loads in the backward) are all issued -- unconditionally, at clamped addresses -- before the first use.  A dropped
// tap keeps the reference's semantics through a select on the loaded value (never a multiply by zero: a NaN at the clamped
// address must not leak).
// grid: (ceil(HW/256), channel slabs, B)
template <int CB>
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
  // clamp the tap offsets so that even dropped taps form a valid address (loaded, then discarded by the select)
  const int onw = vnw ? t.y0 * W + t.x0 : 0, one = vne ? t.y0 * W + t.x0 + 1 : 0;
  const int osw = vsw ? (t.y0 + 1) * W + t.x0 : 0, ose = vse ? (t.y0 + 1) * W + t.x0 + 1 : 0;
  const int c_begin = blockIdx.y * slab, c_end = min(C, c_begin + slab);
  const float* ip = img + ((size_t)b * C + c_begin) * HW;
  // per-thread pointers of the four taps; the channel step (HW floats) is warp-uniform
  const float* pnw = ip + onw;
  const float* pne = ip + one;
  const float* psw = ip + osw;
  const float* pse = ip + ose;
  float* op = out + ((size_t)b * C + c_begin) * HW + pix;
  const size_t step = (size_t)CB * HW;
  for (int c0 = c_begin; c0 < c_end; c0 += CB, pnw += step, pne += step, psw += step, pse += step, op += step) {
    float a[CB][4];
    const int last = c_end - 1 - c0;   // channels past the slab re-load its last channel (discarded below)
#pragma unroll
    for (int j = 0; j < CB; ++j) {
      const unsigned co = (unsigned)(min(j, last) * HW);   // warp-uniform
      a[j][0] = __ldg(pnw + co); a[j][1] = __ldg(pne + co); a[j][2] = __ldg(psw + co); a[j][3] = __ldg(pse + co);
    }
    // no early exit inside the batch (a branch here makes the compiler sink every channel's loads next to their use, which
    // serialises the load latencies -- seen in the SASS of an earlier version): the tail is a predicated store
#pragma unroll
    for (int j = 0; j < CB; ++j) {
      float r = 0.f;   // ATen's tap order: nw, ne, sw, se
      r = fmaf(vnw ? a[j][0] : 0.f, wnw, r);
      r = fmaf(vne ? a[j][1] : 0.f, wne, r);
      r = fmaf(vsw ? a[j][2] : 0.f, wsw, r);
      r = fmaf(vse ? a[j][3] : 0.f, wse, r);
      if (j <= last) op[(unsigned)(j * HW)] = r * mul;
    }
  }
}

// (Measured dead end, twice: a "quad" forward -- a thread owns 4 adjacent samples, the 5 + 5 source pixels of a coherent quad fetched
// as 4 x LDG.128 and aligned with a register funnel, channels in batches of 4 -- issues 3.7x fewer instructions and half the L1
// wavefronts per sample, and is 1.4-2.5x SLOWER at every size (8x32x188x620: 104-156 us against 73-78 us for the kernel above;
// 8x128x188x620: 674 vs 265 us): with 4x fewer threads the gather has less memory-level parallelism than the one-sample-per-
// thread form, whose neighbouring lanes share L1 lines anyway.)

// Backward.  d_img is a scatter (red.global.add.f32, zeroed by the entry point); d_flow / d_occ are per-pixel and
// accumulated across channel slabs with one atomic per slab (plain store when there is 1 slab).
// Lanes are 32 horizontally adjacent samples, so for a spatially coherent flow (what the decoders produce: an up-sampled
// coarse field) the four reds of a warp hit runs of consecutive addresses -- measured (tools/red_bench.cu) the L2 atomic
// path then retires ~1.1 T elements/s, against 0.18 T/s for incoherent addresses -- and the east taps of lane l coincide
// with the west taps of lane l+1: that coincidence is detected once per sample (it does not depend on the channel) and the
// two contributions are summed with one shuffle, halving the red operations.
// (Measured: collecting the left-over east taps of lane 31 of a whole channel batch into ONE red -- 17 instead of 32 red
// instructions per 8 channels -- changes nothing (210 vs 209 us at 8x32x188x620, range map likewise): at that size the scatter is
// bound by the DRAM traffic of the zero-fill + read-modify-write of d_img (dram_rd 365 MB for 238 MB of inputs), not by the
// number of red instructions.)
// (Measured dead end: a third, "interior" path for warps whose lanes all have four valid taps but whose merge plan is not the
// regular one -- no validity selects, unpredicated loads, the merges and the two east reds predicated per lane: 46 instead of
// ~105 instructions per sample and channel in the SASS -- changes nothing: 8x32x96x128 on a rough field 53-54 vs 50-51 us, gentle
// field 37.9 vs 37.9 us, 8x32x188x620 204 vs 209 us.  The same probe shows where the time is: without the d_img scatter the call
// takes 25.6 us at 8x32x96x128 whatever the field, i.e. two dependent memory round trips (flow -> taps) per short-lived CTA x 5.2
// generations of CTAs; the scatter adds 12 us on a coherent field and 25-39 us on a rough one.  Instruction count is not it.)
template <int CB>
__global__ void __launch_bounds__(WARP_THREADS, 2)
warp_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ img, const float* __restrict__ flow,
                const float* __restrict__ occ, float* __restrict__ d_img, float* __restrict__ d_flow,
                float* __restrict__ d_occ, int C, int H, int W, int slab, int nslabs, int flags, float scale) {
  const int HW = H * W;
  const int gidx = blockIdx.x * WARP_THREADS + threadIdx.x;
  const bool act = gidx < HW;
  const int pix = act ? gidx : HW - 1;
  const int y = pix / W, x = pix - y * W;
  const int b = blockIdx.z;
  const bool align = flags & OCF_WARP_ALIGN_CORNERS;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;

  const float u = __fmul_rn(__ldg(flow + ((size_t)b * 2) * HW + pix), scale);
  const float v = __fmul_rn(__ldg(flow + ((size_t)b * 2 + 1) * HW + pix), scale);
  const float ix = unnormalize(__fadd_rn((float)x, u), W, max(W - 1, 1), align);
  const float iy = unnormalize(__fadd_rn((float)y, v), H, max(H - 1, 1), align);
  const Taps t = make_taps(ix, iy, H, W);
  float mul = act ? 1.f : 0.f;
  if (flags & OCF_WARP_IS_MASK) mul *= mask_of(t);
  const float occv = occ != nullptr ? __ldg(occ + (size_t)b * HW + pix) : 1.f;
  const float gmul = mul * occv;  // d out / d sample
  const float wnw = t.wx0 * t.wy0, wne = t.wx1 * t.wy0, wsw = t.wx0 * t.wy1, wse = t.wx1 * t.wy1;
  const int o = t.y0 * W + t.x0;
  const bool vnw = act && t.vx0 && t.vy0, vne = act && t.vx1 && t.vy0, vsw = act && t.vx0 && t.vy1, vse = act && t.vx1 && t.vy1;
  // tap offsets: -1 marks a dropped tap in the merge plan; loads / reds go to the clamped copy
  const int onw = vnw ? o : -1, one = vne ? o + 1 : -1, osw = vsw ? o + W : -1, ose = vse ? o + W + 1 : -1;
  // merge plan (channel independent): my west taps absorb the previous lane's east taps when they are the same address
  const int pne = __shfl_up_sync(full, one, 1), pse = __shfl_up_sync(full, ose, 1);
  const bool take_n = lane > 0 && onw >= 0 && pne == onw;
  const bool take_s = lane > 0 && osw >= 0 && pse == osw;
  const bool give_n = __shfl_down_sync(full, (int)take_n, 1) && lane < 31;
  const bool give_s = __shfl_down_sync(full, (int)take_s, 1) && lane < 31;
  const bool rnw = vnw, rne = vne && !give_n, rsw = vsw, rse = vse && !give_s;   // which of my four reds are issued

  const int c_begin = blockIdx.y * slab, c_end = min(C, c_begin + slab);
  // per-thread pointers of the four taps / the sample; the channel step (HW floats) is warp-uniform
  const float* ip = img + ((size_t)b * C + c_begin) * HW;
  const float* pnw = ip + max(onw, 0);
  const float* pne_ = ip + max(one, 0);
  const float* psw = ip + max(osw, 0);
  const float* pse_ = ip + max(ose, 0);
  const float* gp = gout + ((size_t)b * C + c_begin) * HW + pix;
  float* dp = d_img != nullptr ? d_img + ((size_t)b * C + c_begin) * HW : nullptr;
  float* dnw = dp + max(onw, 0);
  float* dne = dp + max(one, 0);
  float* dsw = dp + max(osw, 0);
  float* dse = dp + max(ose, 0);
  float gx = 0.f, gy = 0.f, go = 0.f;
  const bool need_vals = d_flow != nullptr || d_occ != nullptr;
  const size_t step = (size_t)CB * HW;
  // Interior warps of a coherent flow (every lane inside the image with four valid taps, every lane's west taps equal to its
  // left neighbour's east taps) take a path without selects and predicates: 2 reds per lane and channel + 2 from lane 31.
  // That is ~3x fewer instructions per sample and channel (the general path is issue-bound: ncu, profiles/r2_warp_*).
  const bool fast = __all_sync(full, act && vnw && vne && vsw && vse && (lane == 0 || (take_n && take_s))) && d_occ == nullptr;
  for (int c0 = c_begin; c0 < c_end; c0 += CB) {
    float gv[CB], ta[CB], tb[CB], tc[CB], td[CB];
    const int last = c_end - 1 - c0;
#pragma unroll
    for (int j = 0; j < CB; ++j) {
      const unsigned co = (unsigned)(min(j, last) * HW);   // warp-uniform
      gv[j] = __ldg(gp + co);
      if (need_vals) { ta[j] = __ldg(pnw + co); tb[j] = __ldg(pne_ + co); tc[j] = __ldg(psw + co); td[j] = __ldg(pse_ + co); }
    }
    if (fast && last >= CB - 1) {
#pragma unroll
      for (int j = 0; j < CB; ++j) {
        const float g = gv[j] * gmul;
        if (need_vals) {
          gx += g * ((tb[j] - ta[j]) * t.wy0 + (td[j] - tc[j]) * t.wy1);
          gy += g * ((tc[j] - ta[j]) * t.wx0 + (td[j] - tb[j]) * t.wx1);
        }
        if (dp != nullptr) {
          float cnw = g * wnw, csw = g * wsw;
          const float cne = g * wne, cse = g * wse;
          const float pn = __shfl_up_sync(full, cne, 1), ps = __shfl_up_sync(full, cse, 1);
          if (lane > 0) { cnw += pn; csw += ps; }
          const unsigned co = (unsigned)(j * HW);
          atomicAdd(dnw + co, cnw);
          atomicAdd(dsw + co, csw);
          if (lane == 31) { atomicAdd(dne + co, cne); atomicAdd(dse + co, cse); }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < CB; ++j) {
        const bool live = j <= last;   // predication, not a branch: see the forward kernel
        const float graw = (act && live) ? gv[j] : 0.f;
        const float g = graw * gmul;
        if (need_vals) {
          const float a = vnw ? ta[j] : 0.f, bb = vne ? tb[j] : 0.f, cc = vsw ? tc[j] : 0.f, dd = vse ? td[j] : 0.f;
          // ATen grid_sampler_2d_backward: gix -= nw*(iy_se-iy) ; += ne*(iy_sw-iy) ; -= sw*(iy-iy_ne) ; += se*(iy-iy_nw)
          gx += g * ((bb - a) * t.wy0 + (dd - cc) * t.wy1);
          gy += g * ((cc - a) * t.wx0 + (dd - bb) * t.wx1);
          if (d_occ != nullptr) go += graw * mul * (a * wnw + bb * wne + cc * wsw + dd * wse);
        }
        if (dp != nullptr) {
          float cnw = g * wnw, csw = g * wsw;
          const float cne = g * wne, cse = g * wse;
          const float pn = __shfl_up_sync(full, cne, 1), ps = __shfl_up_sync(full, cse, 1);
          if (take_n) cnw += pn;
          if (take_s) csw += ps;
          const unsigned co = (unsigned)(j * HW);
          if (live && rnw) atomicAdd(dnw + co, cnw);
          if (live && rne) atomicAdd(dne + co, cne);
          if (live && rsw) atomicAdd(dsw + co, csw);
          if (live && rse) atomicAdd(dse + co, cse);
        }
      }
    }
    gp += step; pnw += step; pne_ += step; psw += step; pse_ += step;
    if (dp != nullptr) { dnw += step; dne += step; dsw += step; dse += step; }
  }
  if (!act) return;
  // chain: ATen multiplies by (W-1)/2 resp. W/2, the reference's normalisation by 2/max(W-1,1)
  const float mx = (align ? 0.5f * (float)(W - 1) : 0.5f * (float)W) * (2.0f / (float)max(W - 1, 1)) * scale;
  const float my = (align ? 0.5f * (float)(H - 1) : 0.5f * (float)H) * (2.0f / (float)max(H - 1, 1)) * scale;
  if (d_flow != nullptr) {
    float* fx = d_flow + ((size_t)b * 2) * HW + pix;
    if (nslabs == 1) { fx[0] = gx * mx; fx[HW] = gy * my; }
    else { atomicAdd(fx, gx * mx); atomicAdd(fx + HW, gy * my); }
  }
  if (d_occ != nullptr) {
    float* po = d_occ + (size_t)b * HW + pix;
    if (nslabs == 1) *po = go; else atomicAdd(po, go);
  }
}

// ---- range map ----------------------------------------------------------------------------------
// A thread owns PPT samples 256 apart (a warp: PPT runs of 32 consecutive samples); all 2 x PPT flow loads are in flight
// before the first red, and the CTA drains its reds once for 256 x PPT samples (the one-sample-per-thread form spent its time in
// load latency + the drain of 6144 short-lived CTAs: ncu long_scoreboard / drain / lg_throttle, profiles/r2_stream_*).
template <int PPT>
__global__ void __launch_bounds__(256)
range_map_kernel(const float* __restrict__ flow, float* __restrict__ range, int H, int W) {
  const int HW = H * W;
  const int b = blockIdx.y;
  const int base = blockIdx.x * (256 * PPT) + threadIdx.x;
  const float* fu = flow + ((size_t)b * 2) * HW;
  const float* fv = fu + HW;
  float us[PPT], vs[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int pc = min(base + i * 256, HW - 1);
    us[i] = __ldg(fu + pc);
    vs[i] = __ldg(fv + pc);
  }
  float* r = range + (size_t)b * HW;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int pix = base + i * 256;
    const bool active = pix < HW;
    const int pixc = active ? pix : HW - 1;
    const int y = pixc / W, x = pixc - y * W;
    const float ex = __fadd_rn((float)x, us[i]);      // flow_to_warp, model.py:223-241
    const float ey = __fadd_rn((float)y, vs[i]);
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
    const int p10 = __shfl_up_sync(full, o10, 1), p11 = __shfl_up_sync(full, o11, 1);
    const float q10 = __shfl_up_sync(full, w10, 1), q11 = __shfl_up_sync(full, w11, 1);
    const bool take0 = lane > 0 && o00 >= 0 && p10 == o00;
    const bool take1 = lane > 0 && o01 >= 0 && p11 == o01;
    if (take0) w00 += q10;
    if (take1) w01 += q11;
    const bool give0 = __shfl_down_sync(full, (int)take0, 1) && lane < 31;
    const bool give1 = __shfl_down_sync(full, (int)take1, 1) && lane < 31;
    if (o00 >= 0) atomicAdd(r + o00, w00);
    if (o10 >= 0 && !give0) atomicAdd(r + o10, w10);
    if (o01 >= 0) atomicAdd(r + o01, w01);
    if (o11 >= 0 && !give1) atomicAdd(r + o11, w11);
  }
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

// channels per CTA: a multiple of the load batch, small enough that the grid covers the 148 SMs a few times over
int pick_slab(int C, int HW, int B, int cb, int ctas_per_sm = 8) {
  const long long pix_blocks = ((long long)HW + WARP_THREADS - 1) / WARP_THREADS * B;
  int slab = (C + cb - 1) / cb * cb;
  while (slab > cb && pix_blocks * ((C + slab - 1) / slab) < (long long)ctas_per_sm * OCF_SM_COUNT) slab = ((slab / cb + 1) / 2) * cb;
  if (slab > 32) slab = 32 / cb * cb;
  return slab < 1 ? 1 : slab;
}

// developer knob for tuning runs (channels per load batch): OCF_WARP_CB = 10 * forward + backward, e.g. 84
int warp_cb(bool fwd) {
  static const int m = []() { const char* e = getenv("OCF_WARP_CB"); return e ? atoi(e) : 88; }();
  return fwd ? m / 10 : m % 10;
}
int range_ppt() {
  static const int m = []() { const char* e = getenv("OCF_RANGE_PPT"); return e ? atoi(e) : 4; }();
  return m;
}

}  // namespace

extern "C" int ocf_warp_fwd(const float* img, const float* flow, const float* occ, float* out, int B, int C, int H, int W,
                            int flags, float scale, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(img); OCF_REQUIRE_PTR(flow); OCF_REQUIRE_PTR(out);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE((long long)H * W < (1LL << 28) && B <= 65535, OCF_EUNSUPPORTED);   // 8 channel planes are addressed with 32-bit offsets
  OCF_REQUIRE((flags & ~3) == 0, OCF_EUNSUPPORTED);
  const int HW = H * W;
  const int kcb = warp_cb(true);
  const int cb = (C < 2 || kcb == 1) ? 1 : ((C < 8 || kcb == 4) ? 4 : 8);   // C = 3 (images): one batch of 4, the 4th load repeats channel 2
  const int slab = pick_slab(C, HW, B, cb);
  const int nslabs = (C + slab - 1) / slab;
  OCF_REQUIRE(nslabs <= 65535, OCF_EUNSUPPORTED);
  dim3 grid((HW + WARP_THREADS - 1) / WARP_THREADS, nslabs, B);
  cudaStream_t s = ocf_cast_stream(stream);
  if (cb == 8) warp_fwd_kernel<8><<<grid, WARP_THREADS, 0, s>>>(img, flow, occ, out, C, H, W, slab, flags, scale);
  else if (cb == 4) warp_fwd_kernel<4><<<grid, WARP_THREADS, 0, s>>>(img, flow, occ, out, C, H, W, slab, flags, scale);
  else warp_fwd_kernel<1><<<grid, WARP_THREADS, 0, s>>>(img, flow, occ, out, C, H, W, slab, flags, scale);
  return ocf_launch_status();
}

extern "C" int ocf_warp_bwd(const float* grad_out, const float* img, const float* flow, const float* occ, float* d_img,
                            float* d_flow, float* d_occ, int B, int C, int H, int W, int flags, float scale,
                            ocf_stream_t stream) {
  OCF_REQUIRE_PTR(grad_out); OCF_REQUIRE_PTR(img); OCF_REQUIRE_PTR(flow);
  OCF_REQUIRE(d_img != nullptr || d_flow != nullptr || d_occ != nullptr, OCF_ENULL);
  OCF_REQUIRE(d_occ == nullptr || occ != nullptr, OCF_ENULL);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE((long long)H * W < (1LL << 28) && B <= 65535, OCF_EUNSUPPORTED);   // 8 channel planes are addressed with 32-bit offsets
  OCF_REQUIRE((flags & ~3) == 0, OCF_EUNSUPPORTED);
  cudaStream_t s = ocf_cast_stream(stream);
  const int HW = H * W;
  const int kcb = warp_cb(false);
  const int cb = (C < 2 || kcb == 1) ? 1 : ((C < 8 || kcb == 4) ? 4 : 8);   // C = 3 (images): one batch of 4, the 4th load repeats channel 2
  // developer knob (tuning runs): grid target of the backward in CTAs per SM (2 are resident at 128 registers)
  static const int bwd_target = []() { const char* e = getenv("OCF_WARP_BWD_TARGET"); return e ? atoi(e) : 8; }();
  const int slab = pick_slab(C, HW, B, cb, bwd_target);
  const int nslabs = (C + slab - 1) / slab;
  OCF_REQUIRE(nslabs <= 65535, OCF_EUNSUPPORTED);
  cudaError_t e;
  if (nslabs > 1) {
    if (d_flow != nullptr && (e = cudaMemsetAsync(d_flow, 0, sizeof(float) * (size_t)B * 2 * HW, s)) != cudaSuccess) return (int)e;
    if (d_occ != nullptr && (e = cudaMemsetAsync(d_occ, 0, sizeof(float) * (size_t)B * HW, s)) != cudaSuccess) return (int)e;
  }
  // (Measured dead end: processing the batch in L2-sized chunks -- zero a chunk, scatter into it while it is resident, to save
  // the write-back + re-fetch of the zeroed lines of a d_img larger than the L2 -- was 40 % slower at 119 MB: the chunk grids
  // no longer fill the GPU and every chunk pays two launch gaps.)
  if (d_img != nullptr && (e = cudaMemsetAsync(d_img, 0, sizeof(float) * (size_t)B * C * HW, s)) != cudaSuccess) return (int)e;
  dim3 grid((HW + WARP_THREADS - 1) / WARP_THREADS, nslabs, B);
  if (cb == 8) warp_bwd_kernel<8><<<grid, WARP_THREADS, 0, s>>>(grad_out, img, flow, occ, d_img, d_flow, d_occ, C, H, W, slab, nslabs, flags, scale);
  else if (cb == 4) warp_bwd_kernel<4><<<grid, WARP_THREADS, 0, s>>>(grad_out, img, flow, occ, d_img, d_flow, d_occ, C, H, W, slab, nslabs, flags, scale);
  else warp_bwd_kernel<1><<<grid, WARP_THREADS, 0, s>>>(grad_out, img, flow, occ, d_img, d_flow, d_occ, C, H, W, slab, nslabs, flags, scale);
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
  const int ppt = ((long long)HW * B >= 8LL * 256 * OCF_SM_COUNT * 4) ? range_ppt() : 1;   // small maps: one sample per thread fills the SMs better
  if (ppt >= 8) range_map_kernel<8><<<dim3((HW + 2047) / 2048, B), 256, 0, s>>>(flow, range_out, H, W);
  else if (ppt >= 4) range_map_kernel<4><<<dim3((HW + 1023) / 1024, B), 256, 0, s>>>(flow, range_out, H, W);
  else if (ppt >= 2) range_map_kernel<2><<<dim3((HW + 511) / 512, B), 256, 0, s>>>(flow, range_out, H, W);
  else range_map_kernel<1><<<dim3((HW + 255) / 256, B), 256, 0, s>>>(flow, range_out, H, W);
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
