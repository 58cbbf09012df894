// Developer micro-benchmark (not part of the library): throughput of global / shared floating-point reductions on B200, the
// primitive behind the bilinear scatter of warp backward and the range map (csrc/warp.cu).
//   red.global.add.f32            lanes -> consecutive addresses / lanes -> random addresses
//   red.global.add.v2.f32 / .v4.f32   (sm_90+ vector reductions)     consecutive / random 16-byte slots
//   red.shared.add.f32            random addresses inside a 32 KB tile
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/red_bench tools/red_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int ITERS = 64;

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ void red1(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red2(float* p, float v) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %1};" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red4(float* p, float v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(v) : "memory");
}

// mode 0: scalar, consecutive   1: scalar, random   2: v2 consecutive   3: v2 random   4: v4 consecutive   5: v4 random
// 6: scalar, "bilinear" pattern: 4 reds per lane at (x, x+1, x+W, x+W+1) with lanes on consecutive x
template <int MODE>
__global__ void red_kernel(float* buf, unsigned nslots) {
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned nthreads = gridDim.x * blockDim.x;
#pragma unroll 4
  for (int it = 0; it < ITERS; ++it) {
    const unsigned lin = (unsigned)it * nthreads + tid;
    if (MODE == 0) red1(buf + lin % nslots, 1.f);
    if (MODE == 1) red1(buf + hash32(lin) % nslots, 1.f);
    if (MODE == 2) red2(buf + 2 * (lin % (nslots / 2)), 1.f);
    if (MODE == 3) red2(buf + 2 * (hash32(lin) % (nslots / 2)), 1.f);
    if (MODE == 4) red4(buf + 4 * (lin % (nslots / 4)), 1.f);
    if (MODE == 5) red4(buf + 4 * (hash32(lin) % (nslots / 4)), 1.f);
    if (MODE == 6) {
      const unsigned base = (lin + (hash32(lin >> 5) & 1023u)) % (nslots - 520);
      red1(buf + base, 1.f); red1(buf + base + 1, 1.f); red1(buf + base + 512, 1.f); red1(buf + base + 513, 1.f);
    }
  }
}

// shared-memory reductions, random addresses in an 8 K-float tile; mode 0: red.shared.add.f32, 1: plain read-modify-write (no atomicity;
// the non-atomic floor)
template <int MODE>
__global__ void sred_kernel(float* out) {
  __shared__ float tile[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) tile[i] = 0.f;
  __syncthreads();
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 4
  for (int it = 0; it < ITERS * 4; ++it) {
    const unsigned a = hash32((unsigned)it * 7919u + tid) & 8191u;
    if (MODE == 0) asm volatile("red.shared.add.f32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(tile + a)), "f"(1.f) : "memory");
    else tile[a] += 1.f;
  }
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s += tile[i];
  if (s == -1.f) out[tid] = s;
}

template <typename F>
static float time_us(F launch, int reps = 5) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best * 1e3f;
}

int main() {
  const unsigned nslots = 12u << 18;  // 12.6 MB of floats: the d(img) buffer of pyramid level 2 at B = 8
  float* buf;
  cudaMalloc(&buf, (size_t)nslots * 4);
  cudaMemset(buf, 0, (size_t)nslots * 4);
  const int grid = 148 * 8, block = 256;
  const double lanes = (double)grid * block * ITERS;
  const char* names[7] = {"red.f32 consecutive", "red.f32 random", "red.v2.f32 consecutive", "red.v2.f32 random", "red.v4.f32 consecutive",
                          "red.v4.f32 random", "red.f32 bilinear 4-tap pattern"};
  const int elems[7] = {1, 1, 2, 2, 4, 4, 4};
  for (int m = 0; m < 7; ++m) {
    float us = 0.f;
    switch (m) {
      case 0: us = time_us([&] { red_kernel<0><<<grid, block>>>(buf, nslots); }); break;
      case 1: us = time_us([&] { red_kernel<1><<<grid, block>>>(buf, nslots); }); break;
      case 2: us = time_us([&] { red_kernel<2><<<grid, block>>>(buf, nslots); }); break;
      case 3: us = time_us([&] { red_kernel<3><<<grid, block>>>(buf, nslots); }); break;
      case 4: us = time_us([&] { red_kernel<4><<<grid, block>>>(buf, nslots); }); break;
      case 5: us = time_us([&] { red_kernel<5><<<grid, block>>>(buf, nslots); }); break;
      case 6: us = time_us([&] { red_kernel<6><<<grid, block>>>(buf, nslots); }); break;
    }
    const double instr_lanes = lanes * (m == 6 ? 4 : 1);
    printf("%-34s %9.1f us   %7.2f G lane-ops/s   %7.2f G elements/s   %6.3f cyc/lane/SM @1.965GHz\n", names[m], us, instr_lanes / us * 1e-3,
           instr_lanes * (m == 6 ? 1 : elems[m]) / us * 1e-3, us * 1e-6 * 1.965e9 * 148 / instr_lanes);
  }
  float* out;
  cudaMalloc(&out, (size_t)grid * block * 4);
  for (int m = 0; m < 2; ++m) {
    float us = m == 0 ? time_us([&] { sred_kernel<0><<<grid, block>>>(out); }) : time_us([&] { sred_kernel<1><<<grid, block>>>(out); });
    const double l = lanes * 4;
    printf("%-34s %9.1f us   %7.2f G lane-ops/s   %6.3f cyc/lane/SM\n", m == 0 ? "red.shared.add.f32 random" : "shared += (non-atomic floor)", us, l / us * 1e-3,
           us * 1e-6 * 1.965e9 * 148 / l);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
