// Developer probe (GPU box): one tcgen05.mma (kind::tf32, M=128, N=192, K=8) on MN-major operands in the no-swizzle
// ("interleave") canonical layout, accumulator read back with tcgen05.ld -- pins the shared-memory descriptor fields,
// the instruction descriptor and the TMEM lane/column mapping that csrc/corr_tc.cu relies on, and what the tensor core
// does with the low 13 mantissa bits of an fp32 operand (truncate or round).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tc_probe tools/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 192, K = 8;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long make_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes) {
  unsigned long long d = 0;
  d |= (unsigned long long)((addr >> 4) & 0x3FFF);
  d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  return d;         // layout type 0 = no swizzle, base offset 0
}

// a[m][k], b[n][k] row-major in global; layout: element (mn, k) at (mn / 4) * chunk_stride + k * 16 + (mn % 4) * 4 bytes
__global__ void __launch_bounds__(160) probe(const float* a, const float* b, float* d, unsigned lbo, unsigned sbo, int mn_major) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ unsigned tmem_base;
  float* sa = reinterpret_cast<float*>(smem);
  float* sb = reinterpret_cast<float*>(smem + 8192);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * K; i += blockDim.x) {
    const int m = i / K, k = i % K;
    const int off = mn_major ? (m / 4) * 32 + k * 4 + (m % 4)   // floats: chunk stride 128 B = 32 floats, k stride 16 B = 4 floats
                             : (m / 8) * 64 + (k / 4) * 32 + (m % 8) * 4 + (k % 4);  // K-major interleave: 8 rows x 16 B core matrices
    sa[off] = a[i];
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    const int off = mn_major ? (n / 4) * 32 + k * 4 + (n % 4) : (n / 8) * 64 + (k / 4) * 32 + (n % 8) * 4 + (k % 4);
    sb[off] = b[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tb = tmem_base;
  if (warp == 4 && (tid & 31) == 0) {
    const unsigned long long da = make_desc(smem_u32(sa), lbo, sbo), db = make_desc(smem_u32(sb), lbo, sbo);
    unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    if (mn_major) idesc |= (1u << 15) | (1u << 16);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  if (warp < 4) {
    unsigned ok = 0, spins = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      if (++spins > (1u << 24)) { if ((tid & 31) == 0) printf("probe: timeout waiting for the MMA\n"); break; }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
      unsigned r[8];
      const unsigned taddr = tb + ((unsigned)(warp * 32) << 16) + (unsigned)c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) d[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tb));
}

int main() {
  float *ha = new float[M * K], *hb = new float[N * K], *hd = new float[M * N];
  float *da, *db, *dd;
  CK(cudaMalloc(&da, M * K * 4)); CK(cudaMalloc(&db, N * K * 4)); CK(cudaMalloc(&dd, M * N * 4));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  struct V { unsigned lbo, sbo; int mn; const char* what; };
  const V variants[] = {
      {4096, 128, 1, "MN-major: SBO=128 (16-byte chunks of 4 MN elements, 8 K rows of 16 B each), LBO=K-group stride"},
      {128, 4096, 1, "MN-major: LBO=128 / SBO=K-group stride (roles swapped)"},
      {128, 256, 0, "K-major no-swizzle: LBO=128 (K core step), SBO=256 (8-row step)"},
      {256, 128, 0, "K-major no-swizzle: roles swapped"},
  };
  for (int pass = 0; pass < 2; ++pass) {
    // pass 0: small integers (exact in tf32) -> checks the layout ; pass 1: values with low mantissa bits -> truncate or round?
    for (int i = 0; i < M * K; ++i) { const int m = i / K, k = i % K; ha[i] = pass == 0 ? (float)((m * 3 + k * 5) % 11 - 5) : 1.0f + (float)((m * 7 + k) % 4096 + 1) / 8388608.0f * 1024.0f; }
    for (int i = 0; i < N * K; ++i) { const int n = i / K, k = i % K; hb[i] = pass == 0 ? (float)((n * 7 + k * 3) % 13 - 6) : (k == 0 ? 1.0f : 0.0f); }
    if (pass == 1) for (int i = 0; i < M * K; ++i) if (i % K != 0) ha[i] = 0.f;
    CK(cudaMemcpy(da, ha, M * K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb, N * K * 4, cudaMemcpyHostToDevice));
    for (const V& v : variants) {
      CK(cudaMemset(dd, 0xff, M * N * 4));
      probe<<<1, 160, 65536>>>(da, db, dd, v.lbo, v.sbo, v.mn);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant [%s]: CUDA error %s\n", v.what, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(hd, dd, M * N * 4, cudaMemcpyDeviceToHost));
      if (pass == 0) {
        int bad = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
          float ref = 0.f;
          for (int k = 0; k < K; ++k) ref += ha[m * K + k] * hb[n * K + k];
          if (hd[m * N + n] != ref) { if (bad < 3) printf("   mismatch m=%d n=%d got %g want %g\n", m, n, hd[m * N + n], ref); ++bad; }
        }
        printf("pass 0 variant [%s]: %d / %d mismatches\n", v.what, bad, M * N);
      } else {
        // D[m][0] = a[m][0] * 1: compare with truncation and with round-to-nearest-even to 10 mantissa bits
        int trunc_ok = 0, rna_ok = 0, exact = 0;
        for (int m = 0; m < M; ++m) {
          unsigned u; memcpy(&u, &ha[m * K], 4);
          unsigned t = u & 0xFFFFE000u, r = (u + 0x1000u) & 0xFFFFE000u;
          float ft, fr; memcpy(&ft, &t, 4); memcpy(&fr, &r, 4);
          trunc_ok += hd[m * N] == ft; rna_ok += hd[m * N] == fr; exact += hd[m * N] == ha[m * K];
        }
        printf("pass 1 variant [%s]: of %d operands the product equals trunc(x): %d, round-half-up(x): %d, x itself: %d\n", v.what, M, trunc_ok, rna_ok, exact);
      }
    }
  }
  return 0;
}
