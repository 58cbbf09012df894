// Internal launchers of the tensor-core correlation kernels (corr_tc.cu), called from the C-ABI entry points in corr.cu / level.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

// d = 4 forward on tcgen05 (3xTF32).  norm: optional device pointer to {mean, inv_std} applied to both inputs on load;
// f1n_out: optional destination of the normalised f1 (batch stride f1n_bstride floats, 0 = dense); f2n_out: optional dense
// destination of the normalised f2.
int ocf_corr_fwd_tc_launch(const float* f1, const float* f2, float* out, unsigned char* mask_out, const float* norm, float* f1n_out,
                           long long f1n_bstride, float* f2n_out, int B, int C, int H, int W, long long out_bstride, float leaky_slope,
                           cudaStream_t s);

// corr.cu: ocf_corr_bwd with batch strides for the feature operands (elements, 0 = dense)
int ocf_corr_bwd_impl(const float* grad_out, const float* out_act, const float* f1, const float* f2, float* df1, float* df2, int B, int C,
                      int H, int W, int d, long long g_bstride, long long act_bstride, float leaky_slope, const unsigned char* mask,
                      long long f1_bstride, long long f2_bstride, void* stream);

// corr.cu: (cached) 4-D tensor map of an NCHW fp32 tensor with a {bw, bh, cc, 1} box, zero fill outside; bstride = batch stride in
// elements (0 = dense).  false when the driver entry point is unavailable or the geometry cannot be described.
bool ocf_make_tensor_map(CUtensorMap* map, const float* base, int B, int C, int H, int W, int bw, int bh, int cc, long long bstride);
