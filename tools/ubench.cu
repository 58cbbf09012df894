// Developer micro-benchmarks for the design choices in csrc/corr.cu (not part of the library):
//   * FFMA (3-register) vs fma.rn.f32x2 issue throughput per SM
//   * LDS.128 cost for different lane->address patterns (distinct / shared inside a quarter-warp / shared across quarters)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

__global__ void ffma_kernel(float* out, float a, float b, long long* cyc) {
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], a, b + acc[(i + 7) & 31] * 0.f);
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// plain 3-source FFMA: acc = x*y + acc with x, y registers that change slowly
__global__ void ffma3_kernel(float* out, const float* in, long long* cyc) {
  float acc[64];
  float x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = in[threadIdx.x + i]; y[i] = in[threadIdx.x + 8 + i]; }
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i * 8 + j] = fmaf(x[i], y[j], acc[i * 8 + j]);
    x[it & 7] += 1.0f;  // keep the compiler from hoisting
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long d, a, b;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

__global__ void ffma2_kernel(float* out, const float* in, long long* cyc) {
  unsigned long long acc[32];
  unsigned long long x[4], y[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a = in[threadIdx.x + 2 * i], b = in[threadIdx.x + 2 * i + 1];
    asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(a), "f"(b));
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a = in[threadIdx.x + 8 + i];
    asm("mov.b64 %0, {%1, %1};" : "=l"(y[i]) : "f"(a));
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0ull;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i * 8 + j]) : "l"(x[i]), "l"(y[j]));
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[i]));
    s += a + b;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// LDS.128 patterns.  mode 0: 32 distinct consecutive float4 (512 B); 1: lane/4 (8 unique, sharers adjacent, inside a
// quarter-warp there are 2 unique); 2: lane%8 (8 unique, each quarter-warp reads the same 8); 3: all lanes same; 4: lane/2
__global__ void lds_kernel(float* out, int mode, long long* cyc) {
  __shared__ __align__(16) float sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int idx;
  switch (mode) {
    case 0: idx = lane; break;
    case 1: idx = lane / 4; break;
    case 2: idx = lane % 8; break;
    case 3: idx = 0; break;
    case 4: idx = lane / 2; break;
    case 5: idx = (lane % 8) * 9; break;   // 8 unique, stride 9 float4 = 36 floats (the kernel's row stride)
    default: idx = lane; break;
  }
  const float4* p = reinterpret_cast<const float4*>(sm) + idx;
  float4 acc = make_float4(0, 0, 0, 0);
  unsigned base[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) base[k] = (unsigned)__cvta_generic_to_shared(p) + k * 512;
  long long t0 = clock64();
  for (int it = 0; it < ITERS * 2; ++it) {
    // 4 independent dependent-chains per thread; the loaded words are all zero, so the addresses never change
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned x, y, z, w;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(base[k]) : "memory");
      base[k] += (x | y | z | w);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float *out, *in;
  long long* cyc;
  cudaMalloc(&out, 1 << 24);
  cudaMalloc(&in, 1 << 20);
  cudaMemset(in, 0, 1 << 20);
  cudaMallocManaged(&cyc, 4096 * 8);
  for (int threads : {128, 256, 512}) {
    ffma_kernel<<<148, threads>>>(out, 1.0001f, 0.5f, cyc);
    cudaDeviceSynchronize();
    printf("ffma(2src+imm-ish) threads=%d: %.3f warp-FFMA/clk/SM\n", threads, (double)ITERS * 32 * (threads / 32) / cyc[0]);
    ffma3_kernel<<<148, threads>>>(out, in, cyc);
    cudaDeviceSynchronize();
    printf("ffma3 threads=%d: %.3f warp-FFMA/clk/SM\n", threads, (double)ITERS * 64 * (threads / 32) / cyc[0]);
    ffma2_kernel<<<148, threads>>>(out, in, cyc);
    cudaDeviceSynchronize();
    printf("ffma2 threads=%d: %.3f warp-FFMA2/clk/SM (x2 FMAs each)\n", threads, (double)ITERS * 32 * (threads / 32) / cyc[0]);
  }
  for (int mode = 0; mode <= 5; ++mode)
    for (int threads : {128, 512, 1024}) {
      lds_kernel<<<148, threads>>>(out, mode, cyc);
      cudaDeviceSynchronize();
      printf("lds128 mode=%d threads=%d: %.3f clk per warp-LDS.128 (SM-wide)\n", mode, threads,
             (double)cyc[0] / ((double)ITERS * 8 * (threads / 32)));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
