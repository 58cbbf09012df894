// Bias + LeakyReLU epilogue of the FlowNetCV convolution blocks (nn.Sequential(Conv2d(bias=True), LeakyReLU(0.1)),
// models/networks/cost_volume_flow_net.py:11-15).  PyTorch runs a cuDNN convolution, a bias-add kernel and a LeakyReLU kernel
// (two extra passes over the output), and in the backward a leaky_relu_backward kernel plus a separate reduction for the bias
// gradient.  Here: forward y = lrelu(x + bias[c]) in ONE in-place pass, backward dx = g * lrelu'(y) and dbias[c] = sum dx in ONE
// pass.  Same arithmetic as the three ATen ops ((x + b), then the select); the sign of y equals the sign of x + b for slope > 0.
#include "common.cuh"

namespace {

constexpr int GT = 256;
constexpr int GU = 4;   // float4 groups per thread, all loads in flight

// grid: (chunks, B * C); one (b, c) plane per blockIdx.y
__global__ void __launch_bounds__(GT)
bias_lrelu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ bias, float* __restrict__ y, int C, size_t HW, float slope,
                      bool vec) {
  const int plane = blockIdx.y;
  const float b = __ldg(bias + plane % C);
  const float* xp = x + (size_t)plane * HW;
  float* yp = y + (size_t)plane * HW;
  if (vec) {
    const size_t n4 = HW / 4;
    const size_t base = (size_t)blockIdx.x * (GT * GU) + threadIdx.x;
    float4 v[GU];
#pragma unroll
    for (int k = 0; k < GU; ++k) {
      const size_t i = base + (size_t)k * GT;
      if (i < n4) v[k] = reinterpret_cast<const float4*>(xp)[i];
    }
#pragma unroll
    for (int k = 0; k < GU; ++k) {
      const size_t i = base + (size_t)k * GT;
      if (i < n4) {
        float4 o = v[k];
        o.x += b; o.y += b; o.z += b; o.w += b;
        o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
        o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
        reinterpret_cast<float4*>(yp)[i] = o;
      }
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * (GT * GU * 4) + threadIdx.x; i < min(HW, (size_t)(blockIdx.x + 1) * (GT * GU * 4)); i += GT) {
      const float o = xp[i] + b;
      yp[i] = o > 0.f ? o : o * slope;
    }
  }
}

__global__ void __launch_bounds__(GT)
bias_lrelu_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y, float* __restrict__ dx, float* __restrict__ dbias, int C,
                      size_t HW, float slope, bool vec) {
  const int plane = blockIdx.y;
  const float* gp = g + (size_t)plane * HW;
  const float* yp = y + (size_t)plane * HW;
  float* dp = dx + (size_t)plane * HW;
  float acc = 0.f;
  if (vec) {
    const size_t n4 = HW / 4;
    const size_t base = (size_t)blockIdx.x * (GT * GU) + threadIdx.x;
    float4 gv[GU], yv[GU];
#pragma unroll
    for (int k = 0; k < GU; ++k) {
      const size_t i = base + (size_t)k * GT;
      if (i < n4) { gv[k] = ocf_ldg_stream4(gp + 4 * i); yv[k] = ocf_ldg_stream4(yp + 4 * i); }
    }
#pragma unroll
    for (int k = 0; k < GU; ++k) {
      const size_t i = base + (size_t)k * GT;
      if (i < n4) {
        float4 o;
        o.x = yv[k].x > 0.f ? gv[k].x : gv[k].x * slope; o.y = yv[k].y > 0.f ? gv[k].y : gv[k].y * slope;
        o.z = yv[k].z > 0.f ? gv[k].z : gv[k].z * slope; o.w = yv[k].w > 0.f ? gv[k].w : gv[k].w * slope;
        reinterpret_cast<float4*>(dp)[i] = o;
        acc += (o.x + o.y) + (o.z + o.w);
      }
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * (GT * GU * 4) + threadIdx.x; i < min(HW, (size_t)(blockIdx.x + 1) * (GT * GU * 4)); i += GT) {
      const float o = yp[i] > 0.f ? gp[i] : gp[i] * slope;
      dp[i] = o;
      acc += o;
    }
  }
  // block sum -> one atomic per CTA
  __shared__ float part[GT / 32];
  acc = ocf_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < GT / 32 ? part[threadIdx.x] : 0.f;
    s = ocf_warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(dbias + plane % C, s);
  }
}

}  // namespace

extern "C" int ocf_bias_lrelu_fwd(const float* x, const float* bias, float* y, int B, int C, long long HW, float slope, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(x); OCF_REQUIRE_PTR(bias); OCF_REQUIRE_PTR(y);
  OCF_REQUIRE(B > 0 && C > 0 && HW > 0, OCF_ESHAPE);
  OCF_REQUIRE((long long)B * C <= 65535, OCF_EUNSUPPORTED);
  const bool vec = (HW % 4 == 0) && ocf_aligned16(x) && ocf_aligned16(y);
  const long long per = (long long)GT * GU * 4;
  dim3 grid((unsigned)((HW + per - 1) / per), B * C);
  bias_lrelu_fwd_kernel<<<grid, GT, 0, ocf_cast_stream(stream)>>>(x, bias, y, C, (size_t)HW, slope, vec);
  return ocf_launch_status();
}

extern "C" int ocf_bias_lrelu_bwd(const float* grad_y, const float* y, float* grad_x, float* grad_bias, int B, int C, long long HW, float slope,
                                  ocf_stream_t stream) {
  OCF_REQUIRE_PTR(grad_y); OCF_REQUIRE_PTR(y); OCF_REQUIRE_PTR(grad_x); OCF_REQUIRE_PTR(grad_bias);
  OCF_REQUIRE(B > 0 && C > 0 && HW > 0, OCF_ESHAPE);
  OCF_REQUIRE((long long)B * C <= 65535, OCF_EUNSUPPORTED);
  cudaStream_t s = ocf_cast_stream(stream);
  cudaError_t e = cudaMemsetAsync(grad_bias, 0, sizeof(float) * C, s);
  if (e != cudaSuccess) return (int)e;
  const bool vec = (HW % 4 == 0) && ocf_aligned16(grad_y) && ocf_aligned16(y) && ocf_aligned16(grad_x);
  const long long per = (long long)GT * GU * 4;
  dim3 grid((unsigned)((HW + per - 1) / per), B * C);
  bias_lrelu_bwd_kernel<<<grid, GT, 0, s>>>(grad_y, y, grad_x, grad_bias, C, (size_t)HW, slope, vec);
  return ocf_launch_status();
}
