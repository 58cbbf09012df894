// On-device input pipeline: uint8 HWC frames + [H,W,2] flow as the loaders deliver them ->
// the network's input layout, in one pass.
//
// Replaces, per sample, the host-side chain of the reference: StaticCenterCrop (models/data/datasets.py:50-55, applied at
// :164-166 and to the flow at :182) -> transforms.ToTensor() (uint8 HWC -> float CHW / 255) ->
// transforms.Normalize(0.5, 0.5) (models/lightning_datamodule.py:20-23) -> torch.cat((img1, img2)) (datasets.py:179) and
// flow.transpose(2, 0, 1) (:185).  Shipping the frames as uint8 makes the host->device copy of a training batch 4x smaller
// for the images; fp32 arithmetic reproduces torchvision's op order exactly: (v / 255 - 0.5) / 0.5.
#include "common.cuh"

namespace {

constexpr int PT = 256;

// one thread = one output pixel of one batch item: 6 image values (+ 2 flow values)
__global__ void __launch_bounds__(PT)
pack_pairs_kernel(const unsigned char* __restrict__ img1, const unsigned char* __restrict__ img2, const float* __restrict__ flow_hw2,
                  float* __restrict__ imgs, float* __restrict__ flow, int H0, int W0, int H, int W, int y0, int x0) {
  const int p = blockIdx.x * PT + threadIdx.x;
  if (p >= H * W) return;
  const int b = blockIdx.y;
  const int y = p / W, x = p - y * W;
  const size_t src = ((size_t)b * H0 + (y0 + y)) * W0 + (x0 + x);
  const size_t HW = (size_t)H * W;
  float* o = imgs + (size_t)b * 6 * HW + p;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float a = __fdiv_rn((float)img1[src * 3 + c], 255.0f), d = __fdiv_rn((float)img2[src * 3 + c], 255.0f);
    o[c * HW] = __fdiv_rn(__fsub_rn(a, 0.5f), 0.5f);
    o[(3 + c) * HW] = __fdiv_rn(__fsub_rn(d, 0.5f), 0.5f);
  }
  if (flow_hw2 != nullptr) {
    const float2 f = *reinterpret_cast<const float2*>(flow_hw2 + src * 2);
    float* fo = flow + (size_t)b * 2 * HW + p;
    fo[0] = f.x;
    fo[HW] = f.y;
  }
}

// FlyingChairs2 occlusion mask (models/data/datasets.py:660-669): uint8 [H0,W0] as decoded -> crop -> float -> ToTensor (no
// scaling for a float array) -> occ[occ > 0.5] = 1 ; occ[occ != 1] = 0, i.e. 1.0 where the decoded value is non-zero
__global__ void __launch_bounds__(PT)
pack_occ_kernel(const unsigned char* __restrict__ occ_u8, float* __restrict__ occ, int H0, int W0, int H, int W, int y0, int x0) {
  const int p = blockIdx.x * PT + threadIdx.x;
  if (p >= H * W) return;
  const int b = blockIdx.y;
  const int y = p / W, x = p - y * W;
  const float v = (float)occ_u8[((size_t)b * H0 + (y0 + y)) * W0 + (x0 + x)];
  occ[(size_t)b * H * W + p] = v > 0.5f ? 1.0f : 0.0f;
}

}  // namespace

extern "C" int ocf_pack_occ(const unsigned char* occ_u8, float* occ, int B, int H0, int W0, int H, int W, int y0, int x0,
                            ocf_stream_t stream) {
  OCF_REQUIRE_PTR(occ_u8); OCF_REQUIRE_PTR(occ);
  OCF_REQUIRE(B > 0 && H0 > 0 && W0 > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(y0 >= 0 && x0 >= 0 && y0 + H <= H0 && x0 + W <= W0, OCF_ESHAPE);
  OCF_REQUIRE(B <= 65535, OCF_EUNSUPPORTED);
  dim3 grid((H * W + PT - 1) / PT, B);
  pack_occ_kernel<<<grid, PT, 0, ocf_cast_stream(stream)>>>(occ_u8, occ, H0, W0, H, W, y0, x0);
  return ocf_launch_status();
}

extern "C" int ocf_pack_pairs(const unsigned char* img1, const unsigned char* img2, const float* flow_hw2, float* imgs, float* flow,
                              int B, int H0, int W0, int H, int W, int y0, int x0, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(img1); OCF_REQUIRE_PTR(img2); OCF_REQUIRE_PTR(imgs);
  OCF_REQUIRE(flow_hw2 == nullptr || flow != nullptr, OCF_ENULL);
  OCF_REQUIRE(B > 0 && H0 > 0 && W0 > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(y0 >= 0 && x0 >= 0 && y0 + H <= H0 && x0 + W <= W0, OCF_ESHAPE);
  OCF_REQUIRE(B <= 65535, OCF_EUNSUPPORTED);
  OCF_REQUIRE(flow_hw2 == nullptr || (reinterpret_cast<uintptr_t>(flow_hw2) & 7u) == 0, OCF_EALIGN);
  dim3 grid((H * W + PT - 1) / PT, B);
  pack_pairs_kernel<<<grid, PT, 0, ocf_cast_stream(stream)>>>(img1, img2, flow_hw2, imgs, flow, H0, W0, H, W, y0, x0);
  return ocf_launch_status();
}
