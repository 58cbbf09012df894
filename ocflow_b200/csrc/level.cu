// Fused pyramid level of FlowNetCV (reference models/networks/cost_volume_flow_net.py:171-173, 186-190, ...):
//
//     warp2 = warp(c2, up_flow * scale); c1n, c2n = normalize_features([c1, warp2]); corr = LeakyReLU(cost_volume(c1n, c2n))
//     x = cat(corr, c1n, up_flow, up_feat)
//
// The reference (and round 1 of this repository) materialises warp2, reads it again for the statistics, writes both
// normalised tensors, reads them in the correlation and copies corr + c1n once more in torch.cat.  Here:
//     ocf_warp_fwd            warp2 (raw)
//     ocf_normalize_stats     one pass over c1 and warp2 -> the scalar {mean, inv_std}        (normalize.cu)
//     ocf_level_corr_fwd      tensor-core correlation that normalises its operands ON LOAD (zero padding stays zero after
//                             normalisation, as in the reference), applies 1/C + LeakyReLU, and writes corr AND c1n straight
//                             into the decoder's concat buffer (batch strides) plus c2n for the backward      (corr_tc.cu)
// i.e. 3 launches, no normalised temporaries re-read, no concat copy of the 81 + C widest channels.
// ocf_level_corr_bwd is the correlation backward with the normalised first feature map read in place from that buffer.
#include "common.cuh"
#include "corr_tc.cuh"

extern "C" int ocf_level_corr_fwd(const float* f1, const float* f2, const float* norm, float* out, long long out_bstride, float* f1n_out,
                                  long long f1n_bstride, float* f2n_out, unsigned char* mask_out, int B, int C, int H, int W,
                                  float leaky_slope, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(f1); OCF_REQUIRE_PTR(f2); OCF_REQUIRE_PTR(out);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(out_bstride == 0 || out_bstride >= 81LL * H * W, OCF_ESHAPE);
  OCF_REQUIRE(f1n_bstride == 0 || f1n_bstride >= (long long)C * H * W, OCF_ESHAPE);
  OCF_REQUIRE((f1n_out == nullptr && f2n_out == nullptr) || norm != nullptr, OCF_ENULL);
  return ocf_corr_fwd_tc_launch(f1, f2, out, mask_out, norm, f1n_out, f1n_bstride, f2n_out, B, C, H, W, out_bstride, leaky_slope,
                                ocf_cast_stream(stream));
}

extern "C" int ocf_level_corr_bwd(const float* grad_out, long long g_bstride, const unsigned char* mask, const float* f1n,
                                  long long f1n_bstride, const float* f2n, float* df1, float* df2, int B, int C, int H, int W,
                                  float leaky_slope, ocf_stream_t stream) {
  OCF_REQUIRE(f1n_bstride == 0 || f1n_bstride >= (long long)C * H * W, OCF_ESHAPE);
  return ocf_corr_bwd_impl(grad_out, nullptr, f1n, f2n, df1, df2, B, C, H, W, 4, g_bstride, 0, leaky_slope, mask, f1n_bstride, 0, stream);
}
