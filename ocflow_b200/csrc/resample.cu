// Bilinear resampling with align_corners=True, the glue either side of the hot path:
//   flow1 = F.interpolate(flow2, scale_factor=4, mode='bilinear', align_corners=True) * 20     cost_volume_flow_net.py:245
//   img1_l2 = F.interpolate(img1, scale_factor=0.25, mode='bilinear', align_corners=True)       models/model.py:396
// Same arithmetic as ATen's upsample_bilinear2d (UpSample.h area_pixel_compute_scale / compute_source_index, fp32):
//   r = (in - 1) / (out - 1)   (0 when out == 1) ;  src = r * dst ;  i0 = (int)src ;  i1 = i0 + (i0 < in - 1) ;  l1 = src - i0 ;  l0 = 1 - l1
//   out = l0y * (l0x * v00 + l1x * v01) + l1y * (l0x * v10 + l1x * v11), then the fused scalar multiplier (the reference's "* 20").
// The backward is a GATHER over the output pixels whose two source rows / columns include the input pixel (no atomics, no
// zero-fill, bit-reproducible): ATen scatters with atomicAdd.
#include "common.cuh"

namespace {

__device__ __forceinline__ float scale_of(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

struct Src {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Src source(float r, int dst, int in) {
  Src s;
  const float f = r * (float)dst;
  s.i0 = min((int)f, in - 1);
  s.i1 = s.i0 + (s.i0 < in - 1 ? 1 : 0);
  s.l1 = f - (float)s.i0;
  s.l0 = 1.0f - s.l1;
  return s;
}

// grid: (ceil(Wo / 4 / 64) ... flat over output groups of 4 pixels, planes) -- a thread produces 4 adjacent output pixels
__global__ void __launch_bounds__(256)
resize_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int Hi, int Wi, int Ho, int Wo, float mul, bool vec) {
  const int plane = blockIdx.y;
  const int W4 = (Wo + 3) >> 2;
  const int gi = blockIdx.x * 256 + threadIdx.x;
  if (gi >= Ho * W4) return;
  const int y = gi / W4, x = (gi - y * W4) << 2;
  const float ry = scale_of(Hi, Ho), rx = scale_of(Wi, Wo);
  const Src sy = source(ry, y, Hi);
  const float* r0 = in + ((size_t)plane * Hi + sy.i0) * Wi;
  const float* r1 = in + ((size_t)plane * Hi + sy.i1) * Wi;
  float res[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int xx = min(x + j, Wo - 1);
    const Src sx = source(rx, xx, Wi);
    const float v00 = __ldg(r0 + sx.i0), v01 = __ldg(r0 + sx.i1), v10 = __ldg(r1 + sx.i0), v11 = __ldg(r1 + sx.i1);
    res[j] = (sy.l0 * (sx.l0 * v00 + sx.l1 * v01) + sy.l1 * (sx.l0 * v10 + sx.l1 * v11)) * mul;
  }
  float* o = out + ((size_t)plane * Ho + y) * Wo + x;
  if (vec && x + 3 < Wo) *reinterpret_cast<float4*>(o) = make_float4(res[0], res[1], res[2], res[3]);
  else {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (x + j < Wo) o[j] = res[j];
  }
}

// d in[y1, x1] = mul * sum over output (y2, x2) of wy(y2 -> y1) * wx(x2 -> x1) * g[y2, x2].
// The output rows that touch input row y1 are those with i0 == y1 (weight l0) or i1 == y1 with i1 != i0 (weight l1): a
// contiguous range around y1 / ry, found by scanning a conservative window and testing membership with the forward's own
// index arithmetic (so forward and backward can never disagree about a rounding).
__global__ void __launch_bounds__(256)
resize_bwd_kernel(const float* __restrict__ g, float* __restrict__ din, int Hi, int Wi, int Ho, int Wo, float mul) {
  const int plane = blockIdx.y;
  const int pi = blockIdx.x * 256 + threadIdx.x;
  if (pi >= Hi * Wi) return;
  const int y1 = pi / Wi, x1 = pi - y1 * Wi;
  const float ry = scale_of(Hi, Ho), rx = scale_of(Wi, Wo);
  // candidate output rows: src = ry * y2 in (y1 - 1, y1 + 1)
  int ylo, yhi, xlo, xhi;
  if (ry > 0.f) { ylo = max((int)floorf((float)(y1 - 1) / ry) - 1, 0); yhi = min((int)ceilf((float)(y1 + 1) / ry) + 1, Ho - 1); }
  else { ylo = 0; yhi = Ho - 1; }
  if (rx > 0.f) { xlo = max((int)floorf((float)(x1 - 1) / rx) - 1, 0); xhi = min((int)ceilf((float)(x1 + 1) / rx) + 1, Wo - 1); }
  else { xlo = 0; xhi = Wo - 1; }
  const float* gp = g + (size_t)plane * Ho * Wo;
  float acc = 0.f;
  for (int y2 = ylo; y2 <= yhi; ++y2) {
    const Src sy = source(ry, y2, Hi);
    float wy = 0.f;
    if (sy.i0 == y1) wy += sy.l0;
    if (sy.i1 == y1) wy += sy.l1;     // i1 == i0 at the last row: both weights land on it, as in the forward
    if (wy == 0.f && sy.i0 != y1 && sy.i1 != y1) continue;
    float row = 0.f;
    for (int x2 = xlo; x2 <= xhi; ++x2) {
      const Src sx = source(rx, x2, Wi);
      float wx = 0.f;
      if (sx.i0 == x1) wx += sx.l0;
      if (sx.i1 == x1) wx += sx.l1;
      if (sx.i0 == x1 || sx.i1 == x1) row = fmaf(wx, __ldg(gp + (size_t)y2 * Wo + x2), row);
    }
    acc = fmaf(wy, row, acc);
  }
  din[(size_t)plane * Hi * Wi + pi] = acc * mul;
}

}  // namespace

extern "C" int ocf_resize_bilinear_fwd(const float* in, float* out, int planes, int Hi, int Wi, int Ho, int Wo, float mul,
                                       ocf_stream_t stream) {
  OCF_REQUIRE_PTR(in); OCF_REQUIRE_PTR(out);
  OCF_REQUIRE(planes > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, OCF_ESHAPE);
  OCF_REQUIRE(planes <= 65535 && (long long)Ho * Wo < (1LL << 30) && (long long)Hi * Wi < (1LL << 30), OCF_EUNSUPPORTED);
  const bool vec = (Wo % 4 == 0) && ocf_aligned16(out);
  const int groups = Ho * ((Wo + 3) >> 2);
  resize_fwd_kernel<<<dim3((groups + 255) / 256, planes), 256, 0, ocf_cast_stream(stream)>>>(in, out, Hi, Wi, Ho, Wo, mul, vec);
  return ocf_launch_status();
}

extern "C" int ocf_resize_bilinear_bwd(const float* grad_out, float* grad_in, int planes, int Hi, int Wi, int Ho, int Wo, float mul,
                                       ocf_stream_t stream) {
  OCF_REQUIRE_PTR(grad_out); OCF_REQUIRE_PTR(grad_in);
  OCF_REQUIRE(planes > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, OCF_ESHAPE);
  OCF_REQUIRE(planes <= 65535 && (long long)Ho * Wo < (1LL << 30) && (long long)Hi * Wi < (1LL << 30), OCF_EUNSUPPORTED);
  resize_bwd_kernel<<<dim3((Hi * Wi + 255) / 256, planes), 256, 0, ocf_cast_stream(stream)>>>(grad_out, grad_in, Hi, Wi, Ho, Wo, mul);
  return ocf_launch_status();
}
