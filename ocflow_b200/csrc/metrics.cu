// End-point-error metrics as one fused reduction, fp32.
//
// Replaces flow_error (reference models/data/utils/flow_utils.py:179-232) and flow_kitti_error (:234-271), which run
// on the host in numpy on [H,W] maps copied back from the GPU: here gt / pred stay on the device ([B,2,H,W] NCHW).
//   plain  (flow_error):        pixels with |gt_u| or |gt_v| > 1e7 are zeroed in all four maps (epe 0, still counted);
//                               sums[0] = sum epe, sums[1] = number of pixels.
//   masked (flow_kitti_error):  only pixels with mask != 0 count; sums[2] = number of outliers
//                               (epe > 3 and epe / (|gt| + 1e-5) > 0.05).
#include "common.cuh"

namespace {

constexpr int MT = 256;

__global__ void __launch_bounds__(MT)
flow_metrics_kernel(const float* __restrict__ gt, const float* __restrict__ pred, const float* __restrict__ mask,
                    double* __restrict__ sums, int HW, size_t npix, int kitti) {
  float acc[3] = {0.f, 0.f, 0.f};
  const size_t stride = (size_t)gridDim.x * MT;
  for (size_t i = (size_t)blockIdx.x * MT + threadIdx.x; i < npix; i += stride) {
    const size_t b = i / HW, p = i - b * HW;
    float tu = __ldg(gt + (b * 2) * HW + p), tv = __ldg(gt + (b * 2 + 1) * HW + p);
    float u = __ldg(pred + (b * 2) * HW + p), v = __ldg(pred + (b * 2 + 1) * HW + p);
    if (!kitti) {
      if (fabsf(tu) > 1e7f || fabsf(tv) > 1e7f) { tu = 0.f; tv = 0.f; u = 0.f; v = 0.f; }
      const float du = tu - u, dv = tv - v;
      acc[0] += sqrtf(du * du + dv * dv);
      acc[1] += 1.f;
    } else {
      if (mask != nullptr && __ldg(mask + i) == 0.f) continue;
      const float du = tu - u, dv = tv - v;
      const float epe = sqrtf(du * du + dv * dv);
      const float mag = sqrtf(tu * tu + tv * tv) + 1e-5f;
      acc[0] += epe;
      acc[1] += 1.f;
      acc[2] += (epe > 3.f && epe / mag > 0.05f) ? 1.f : 0.f;
    }
  }
  ocf_block_accumulate<3>(acc, sums);
}

}  // namespace

extern "C" int ocf_flow_metrics(const float* gt, const float* pred, const float* mask, double* sums, int B, int H, int W, int kitti,
                                ocf_stream_t stream) {
  OCF_REQUIRE_PTR(gt); OCF_REQUIRE_PTR(pred); OCF_REQUIRE_PTR(sums);
  OCF_REQUIRE(B > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(kitti == 0 || kitti == 1, OCF_EUNSUPPORTED);
  OCF_REQUIRE(mask == nullptr || kitti == 1, OCF_EUNSUPPORTED);
  cudaStream_t s = ocf_cast_stream(stream);
  cudaError_t e = cudaMemsetAsync(sums, 0, 3 * sizeof(double), s);
  if (e != cudaSuccess) return (int)e;
  const size_t npix = (size_t)B * H * W;
  size_t blocks = (npix + MT - 1) / MT;
  if (blocks > (size_t)8 * OCF_SM_COUNT) blocks = (size_t)8 * OCF_SM_COUNT;
  flow_metrics_kernel<<<(unsigned)blocks, MT, 0, s>>>(gt, pred, mask, sums, H * W, npix, kitti);
  return ocf_launch_status();
}
