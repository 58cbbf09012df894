// Shared helpers for the ocflow_b200 kernels (sm_100a only; no other architecture is built).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ocflow_b200.h"

#define OCF_SM_COUNT 148  // B200: 2 dies x 74 SMs; grids of the streaming kernels are sized from this

#define OCF_REQUIRE_PTR(p) \
  do {                     \
    if ((p) == nullptr) return OCF_ENULL; \
  } while (0)

#define OCF_REQUIRE(cond, code) \
  do {                          \
    if (!(cond)) return (code); \
  } while (0)

static inline int ocf_launch_status() {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the sticky launch-configuration error so the next call is clean
    return (int)e;
  }
  return OCF_OK;
}

static inline cudaStream_t ocf_cast_stream(ocf_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline bool ocf_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float ocf_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double ocf_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of NV per-thread floats, accumulated into doubles in global memory with one atomic
// per block per value.  blockDim.x must be a multiple of 32 and <= 1024.
template <int NV>
__device__ __forceinline__ void ocf_block_accumulate(const float (&v)[NV], double* dst) {
  __shared__ float part[NV][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float s = ocf_warp_sum(v[i]);
    if (lane == 0) part[i][wid] = s;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = (lane < nw) ? (double)part[i][lane] : 0.0;
      s = ocf_warp_sum(s);
      if (lane == 0) atomicAdd(dst + i, s);
    }
  }
}

// Streaming (read-once) global load that does not allocate in L1.
__device__ __forceinline__ float ocf_ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ float4 ocf_ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// ---- quad gather / scatter helpers (warp.cu, loss.cu) ------------------------------------------------
// 4 horizontally adjacent bilinear samples of a coherent flow touch 5 consecutive source pixels per tap row, i.e. two
// 16-byte aligned groups q[0..7]; o = (first column) & 3 is lane dependent and is resolved with select chains.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// v[j] = q[o + j], j = 0..4, o in 0..3
__device__ __forceinline__ void funnel_gather(const float (&q)[8], int o, float (&v)[5]) {
  float t[6];
  const bool s2 = o & 2, s1 = o & 1;
#pragma unroll
  for (int k = 0; k < 6; ++k) t[k] = s2 ? q[k + 2] : q[k];
#pragma unroll
  for (int j = 0; j < 5; ++j) v[j] = s1 ? t[j + 1] : t[j];
}

// q[k] = v[k - o] (0 outside 0..4), k = 0..7
__device__ __forceinline__ void funnel_scatter(const float (&v)[5], int o, float (&q)[8]) {
  float t[6];
  const bool s2 = o & 2, s1 = o & 1;
#pragma unroll
  for (int k = 0; k < 6; ++k) t[k] = s1 ? (k >= 1 ? v[k - 1] : 0.f) : (k < 5 ? v[k] : 0.f);
#pragma unroll
  for (int k = 0; k < 8; ++k) q[k] = s2 ? (k >= 2 ? t[k - 2] : 0.f) : (k < 6 ? t[k] : 0.f);
}

__device__ __forceinline__ float4 ldg4_or_zero(const float* p, bool ok) {
  return ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

