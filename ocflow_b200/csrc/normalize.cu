// Feature normalisation forward/backward (reference models/networks/correlation_layer.py:42-82).
//
// A "group" is the set of elements one var_mean reduces over: one sample of one tensor
// (moments_across_channels) or one channel of one sample.  With moments_across_images the reference
// averages the per-group means and variances of ALL tensors to a single scalar pair (:66-68) -- the
// variance is the mean of per-group biased variances, not the pooled variance.
//
// forward : (1) partial sums per group (128-bit loads, fp64 atomics); the LAST CTA to finish (ticket counter in the
//           workspace) turns the sums into the applied {mean, inv_std} per group, (2) streaming apply
//           y = (x - mean) * inv_std.  Two launches; the second read of x comes out of the 126 MB L2.
// backward: same two steps with sum(g), sum(g*x): the reference does not detach the statistics, so
//           d x_i = g_i*r + A + Bc*(x_i - mu_group(i)).
//
// Workspace layout (floats, caller-owned; NG = T*B*G groups, G = 1 or C):
//   stats: [0, 4NG)  2NG doubles  sum(x), sum(x^2) per group           (fwd scratch)
//          [4NG,6NG) {mean, var} per group
//          [6NG,8NG) {mean_applied, inv_std_applied} per group
//          [8NG]     ticket counter of the sums kernel (unsigned)
//   red  : [0, 4NG)  2NG doubles  sum(g), sum(g*x) per group           (bwd scratch)
//          [4NG,7NG) {r', A, Bc} per group
//          [7NG]     ticket counter
#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int MAX_T = 8;
constexpr int UB = 8;   // float4 groups per thread and tensor in the vectorised kernels

struct PtrPack {
  const float* in[MAX_T];
  const float* in2[MAX_T];
  float* out[MAX_T];
  long long out_bstride[MAX_T];   // elements between batch items of out[t]; 0 = dense (norm_apply_kernel only)
};

// The forward sums are taken about a per-group pivot (the group's first element): var = E[(x-p)^2] - E[x-p]^2 is shift
// invariant and loses nothing to cancellation when |mean| >> std (the plain E[x^2] - mean^2 form from fp32 partial sums
// does; torch.var_mean, which the reference uses, does not).
__device__ __forceinline__ void finalize_fwd(float* __restrict__ stats, int NG, double inv_len, int flags, const PtrPack& pk, int BG,
                                             size_t glen) {
  const volatile double* sums = reinterpret_cast<const volatile double*>(stats);  // written by other CTAs' atomics
  float* grp = stats + 4 * (size_t)NG;
  float* app = stats + 6 * (size_t)NG;
  __shared__ double sm[2];
  if (threadIdx.x == 0) { sm[0] = 0.0; sm[1] = 0.0; }
  __syncthreads();
  double lm = 0.0, lv = 0.0;
  for (int g = threadIdx.x; g < NG; g += blockDim.x) {
    const int t = g / BG;
    const double pivot = (double)pk.in[t][(size_t)(g - t * BG) * glen];
    const double ms = sums[2 * g] * inv_len;          // mean of (x - pivot)
    double v = sums[2 * g + 1] * inv_len - ms * ms;
    if (v < 0.0) v = 0.0;
    const double m = pivot + ms;
    grp[2 * g] = (float)m;
    grp[2 * g + 1] = (float)v;
    lm += m; lv += v;
  }
  lm = ocf_warp_sum(lm); lv = ocf_warp_sum(lv);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&sm[0], lm); atomicAdd(&sm[1], lv); }
  __syncthreads();
  const bool across = flags & OCF_NORM_ACROSS_IMAGES;
  const float gm = (float)(sm[0] / NG), gv = (float)(sm[1] / NG);
  for (int g = threadIdx.x; g < NG; g += blockDim.x) {
    const float m = across ? gm : grp[2 * g];
    const float v = across ? gv : grp[2 * g + 1];
    app[2 * g] = (flags & OCF_NORM_CENTER) ? m : 0.f;
    app[2 * g + 1] = (flags & OCF_NORM_NORMALIZE) ? 1.0f / sqrtf(v + 1e-16f) : 1.f;
  }
}

__device__ __forceinline__ void finalize_bwd(const float* __restrict__ stats, float* __restrict__ red, int NG, double glen, int flags) {
  const volatile double* sums = reinterpret_cast<const volatile double*>(red);  // sum g, sum g*x per group
  const float* app = stats + 6 * (size_t)NG;
  float* co = red + 4 * (size_t)NG;
  const bool across = flags & OCF_NORM_ACROSS_IMAGES, nz = flags & OCF_NORM_NORMALIZE, ce = flags & OCF_NORM_CENTER;
  __shared__ double sm[2];
  if (threadIdx.x == 0) { sm[0] = 0.0; sm[1] = 0.0; }
  __syncthreads();
  double sg = 0.0, sgx = 0.0;
  for (int g = threadIdx.x; g < NG; g += blockDim.x) { sg += sums[2 * g]; sgx += sums[2 * g + 1]; }
  sg = ocf_warp_sum(sg); sgx = ocf_warp_sum(sgx);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&sm[0], sg); atomicAdd(&sm[1], sgx); }
  __syncthreads();
  for (int g = threadIdx.x; g < NG; g += blockDim.x) {
    const double m = app[2 * g], r = app[2 * g + 1];           // applied mean (0 when !center), inv_std (1 when !normalize)
    const double Sg = across ? sm[0] : sums[2 * g];
    const double Sgx = across ? sm[1] : sums[2 * g + 1];
    const double cnt = across ? glen * NG : glen;               // elements the statistics average over
    // y = (x - m) * r ;  dL/dm = -r*Sg (center) ; dL/dr = Sgx - m*Sg (normalize) ; r = (v+eps)^-1/2
    const double dLdm = ce ? -r * Sg : 0.0;
    const double dLdv = nz ? (Sgx - m * Sg) * (-0.5 * r * r * r) : 0.0;
    co[3 * g] = (float)r;
    co[3 * g + 1] = (float)(dLdm / cnt);
    co[3 * g + 2] = (float)(dLdv * 2.0 / cnt);                  // multiplies (x_i - mu_group(i))
  }
}

// grid: (chunks, NG).  Group gidx = (t*B + b)*G + gc covers `glen` contiguous floats.
// fwd: s0 = sum x, s1 = sum x^2 ; bwd (SECOND: x = grad, w = input): s0 = sum g, s1 = sum g*x.
// The last CTA (ticket) runs the finalize step, so no separate launch is needed between the sums and the apply pass.
template <bool SECOND, bool VEC>
__global__ void __launch_bounds__(NT)
group_sums_kernel(PtrPack pk, int B, int G, size_t glen, float* __restrict__ ws, unsigned* __restrict__ ticket,
                  const float* __restrict__ stats_for_bwd, int NG, int flags) {
  const int gidx = blockIdx.y;
  const int t = gidx / (B * G), rem = gidx - t * (B * G);
  const float* x = pk.in[t] + (size_t)rem * glen;
  const float* w = SECOND ? pk.in2[t] + (size_t)rem * glen : nullptr;
  const float pivot = SECOND ? 0.f : __ldg(x);   // forward: sums of (x - pivot), see finalize_fwd
  float acc[2] = {0.f, 0.f};
  if (VEC) {
    // a CTA owns NT * UB consecutive float4 groups; a thread's UB loads (2 UB in the backward) are all issued before the
    // first use -- the one-load-per-iteration form was latency bound (ncu: long_scoreboard, SMs idle half of the time)
    const size_t n4 = glen / 4;
    const size_t base = (size_t)blockIdx.x * (NT * UB) + threadIdx.x;
    float4 u[UB], p[UB];
#pragma unroll
    for (int k = 0; k < UB; ++k) {
      const size_t i = base + (size_t)k * NT;
      u[k] = i < n4 ? ocf_ldg_stream4(x + 4 * i) : make_float4(pivot, pivot, pivot, pivot);
      if (SECOND) p[k] = i < n4 ? ocf_ldg_stream4(w + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;  // two independent chains per sum
#pragma unroll
    for (int k = 0; k < UB; ++k) {
      float4 v = u[k];
      if (!SECOND) { v.x -= pivot; v.y -= pivot; v.z -= pivot; v.w -= pivot; }
      else if (base + (size_t)k * NT >= n4) v = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 q = SECOND ? p[k] : v;
      if (k & 1) { a1 += (v.x + v.y) + (v.z + v.w); b1 = fmaf(v.x, q.x, fmaf(v.y, q.y, fmaf(v.z, q.z, fmaf(v.w, q.w, b1)))); }
      else { a0 += (v.x + v.y) + (v.z + v.w); b0 = fmaf(v.x, q.x, fmaf(v.y, q.y, fmaf(v.z, q.z, fmaf(v.w, q.w, b0)))); }
    }
    acc[0] = a0 + a1;
    acc[1] = b0 + b1;
  } else {
    const size_t per = (glen + gridDim.x - 1) / gridDim.x;
    const size_t lo = (size_t)blockIdx.x * per, hi = min(glen, lo + per);
    for (size_t i = lo + threadIdx.x; i < hi; i += NT) {
      const float a = x[i] - pivot;
      const float bb = SECOND ? w[i] : a;
      acc[0] += a;
      acc[1] = fmaf(a, bb, acc[1]);
    }
  }
  ocf_block_accumulate<2>(acc, reinterpret_cast<double*>(ws) + 2 * (size_t)gidx);
  // ---- last CTA finalises ----
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (SECOND) finalize_bwd(stats_for_bwd, ws, NG, (double)glen, flags);
  else finalize_fwd(ws, NG, 1.0 / (double)glen, flags, pk, B * G, glen);
}

// y = (x - m) * r per group.  grid: (chunks, NG)
__global__ void __launch_bounds__(NT)
norm_apply_kernel(PtrPack pk, int B, int G, size_t glen, const float* __restrict__ app, bool vec) {
  const int gidx = blockIdx.y;
  const int t = gidx / (B * G), rem = gidx - t * (B * G);
  const float* x = pk.in[t] + (size_t)rem * glen;
  float* y = pk.out[t] + (size_t)rem * glen;
  if (pk.out_bstride[t] != 0) {   // output inside a wider (concat) buffer: only the batch stride differs
    const int bb = rem / G;
    y += (size_t)bb * ((size_t)pk.out_bstride[t] - glen * (size_t)G);
  }
  const float m = app[2 * gidx], r = app[2 * gidx + 1];
  if (vec) {
    const size_t n4 = glen / 4;
    const size_t base = (size_t)blockIdx.x * (NT * UB) + threadIdx.x;
    float4 v[UB];
#pragma unroll
    for (int k = 0; k < UB; ++k) {
      const size_t i = base + (size_t)k * NT;
      if (i < n4) v[k] = ocf_ldg_stream4(x + 4 * i);
    }
#pragma unroll
    for (int k = 0; k < UB; ++k) {
      const size_t i = base + (size_t)k * NT;
      if (i < n4) {
        float4 o = v[k];
        o.x = (o.x - m) * r; o.y = (o.y - m) * r; o.z = (o.z - m) * r; o.w = (o.w - m) * r;
        reinterpret_cast<float4*>(y)[i] = o;
      }
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < glen; i += (size_t)gridDim.x * NT) y[i] = (x[i] - m) * r;
  }
}

__global__ void __launch_bounds__(NT)
norm_bwd_apply_kernel(PtrPack pk, int B, int G, size_t glen, const float* __restrict__ stats, const float* __restrict__ red, int NG,
                      bool vec) {
  const int gidx = blockIdx.y;
  const int t = gidx / (B * G), rem = gidx - t * (B * G);
  const float* g = pk.in[t] + (size_t)rem * glen;
  const float* x = pk.in2[t] + (size_t)rem * glen;
  float* dx = pk.out[t] + (size_t)rem * glen;
  const float mu = stats[4 * (size_t)NG + 2 * gidx];
  const float* co = red + 4 * (size_t)NG + 3 * (size_t)gidx;
  const float r = co[0], A = co[1], Bc = co[2];
  if (vec) {
    const size_t n4 = glen / 4;
    const size_t base = (size_t)blockIdx.x * (NT * UB) + threadIdx.x;
    float4 gv[UB], xv[UB];
#pragma unroll
    for (int k = 0; k < UB; ++k) {
      const size_t i = base + (size_t)k * NT;
      if (i < n4) { gv[k] = ocf_ldg_stream4(g + 4 * i); xv[k] = ocf_ldg_stream4(x + 4 * i); }
    }
#pragma unroll
    for (int k = 0; k < UB; ++k) {
      const size_t i = base + (size_t)k * NT;
      if (i < n4) {
        float4 o;
        o.x = fmaf(gv[k].x, r, fmaf(Bc, xv[k].x - mu, A)); o.y = fmaf(gv[k].y, r, fmaf(Bc, xv[k].y - mu, A));
        o.z = fmaf(gv[k].z, r, fmaf(Bc, xv[k].z - mu, A)); o.w = fmaf(gv[k].w, r, fmaf(Bc, xv[k].w - mu, A));
        reinterpret_cast<float4*>(dx)[i] = o;
      }
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < glen; i += (size_t)gridDim.x * NT)
      dx[i] = fmaf(g[i], r, fmaf(Bc, x[i] - mu, A));
  }
}

// vectorised kernels: a CTA covers NT * UB float4 groups ; scalar fallbacks: ~4 CTAs per SM overall, >= 2048 elements each
int chunks_for(size_t glen, int NG, bool vec) {
  if (vec) {
    const size_t n4 = glen / 4;
    long long c = (long long)((n4 + (size_t)NT * UB - 1) / ((size_t)NT * UB));
    return (int)(c < 1 ? 1 : c);
  }
  long long want = (4LL * OCF_SM_COUNT + NG - 1) / NG;
  long long maxc = (long long)((glen + 2047) / 2048);
  if (want > maxc) want = maxc;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  return (int)want;
}

}  // namespace

// statistics only: fills the stats workspace (per-group {mean, var} and the applied {mean, inv_std}); the apply pass is either
// norm_apply_kernel (ocf_normalize_fwd) or folded into the consumer (ocf_level_corr_fwd normalises on load)
extern "C" int ocf_normalize_stats(const float* const* xs, int T, int B, int C, int H, int W, int flags, float* stats, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(xs); OCF_REQUIRE_PTR(stats);
  OCF_REQUIRE(T > 0 && T <= MAX_T, OCF_EUNSUPPORTED);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE((flags & ~15) == 0, OCF_EUNSUPPORTED);
  OCF_REQUIRE((reinterpret_cast<uintptr_t>(stats) & 7u) == 0, OCF_EALIGN);
  PtrPack pk;
  bool vec = true;
  for (int t = 0; t < T; ++t) {
    OCF_REQUIRE_PTR(xs[t]);
    pk.in[t] = xs[t]; pk.in2[t] = nullptr; pk.out[t] = nullptr; pk.out_bstride[t] = 0;
    vec = vec && ocf_aligned16(xs[t]);
  }
  const int G = (flags & OCF_NORM_ACROSS_CHANNELS) ? 1 : C;
  const size_t glen = (size_t)(G == 1 ? C : 1) * H * W;
  const long long NGll = (long long)T * B * G;
  OCF_REQUIRE(NGll <= 65535, OCF_EUNSUPPORTED);
  const int NG = (int)NGll;
  vec = vec && (glen % 4 == 0);
  cudaStream_t s = ocf_cast_stream(stream);
  unsigned* ticket = reinterpret_cast<unsigned*>(stats + 8 * (size_t)NG);
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(float) * (8 * (size_t)NG + 2), s);  // fp64 sums and the ticket in one go
  if (e != cudaSuccess) return (int)e;
  const int chunks = chunks_for(glen, NG, vec);
  if (vec) group_sums_kernel<false, true><<<dim3(chunks, NG), NT, 0, s>>>(pk, B, G, glen, stats, ticket, nullptr, NG, flags);
  else group_sums_kernel<false, false><<<dim3(chunks, NG), NT, 0, s>>>(pk, B, G, glen, stats, ticket, nullptr, NG, flags);
  return ocf_launch_status();
}

extern "C" int ocf_normalize_fwd(const float* const* xs, float* const* ys, int T, int B, int C, int H, int W, int flags, float* stats,
                                 ocf_stream_t stream) {
  OCF_REQUIRE_PTR(ys);
  if (int st = ocf_normalize_stats(xs, T, B, C, H, W, flags, stats, stream)) return st;
  PtrPack pk;
  bool vec = true;
  for (int t = 0; t < T; ++t) {
    OCF_REQUIRE_PTR(ys[t]);
    pk.in[t] = xs[t]; pk.in2[t] = nullptr; pk.out[t] = ys[t]; pk.out_bstride[t] = 0;
    vec = vec && ocf_aligned16(xs[t]) && ocf_aligned16(ys[t]);
  }
  const int G = (flags & OCF_NORM_ACROSS_CHANNELS) ? 1 : C;
  const size_t glen = (size_t)(G == 1 ? C : 1) * H * W;
  const int NG = T * B * G;
  vec = vec && (glen % 4 == 0);
  norm_apply_kernel<<<dim3(chunks_for(glen, NG, vec), NG), NT, 0, ocf_cast_stream(stream)>>>(pk, B, G, glen, stats + 6 * (size_t)NG, vec);
  return ocf_launch_status();
}

// apply pass alone, with previously computed statistics; ys[t] may live inside a wider buffer (batch stride in elements, 0 = dense)
extern "C" int ocf_normalize_apply(const float* const* xs, float* const* ys, const long long* y_bstrides, int T, int B, int C, int H, int W,
                                   int flags, const float* stats, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(xs); OCF_REQUIRE_PTR(ys); OCF_REQUIRE_PTR(stats);
  OCF_REQUIRE(T > 0 && T <= MAX_T, OCF_EUNSUPPORTED);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE((flags & ~15) == 0, OCF_EUNSUPPORTED);
  PtrPack pk;
  bool vec = true;
  for (int t = 0; t < T; ++t) {
    OCF_REQUIRE_PTR(xs[t]); OCF_REQUIRE_PTR(ys[t]);
    const long long bs = y_bstrides != nullptr ? y_bstrides[t] : 0;
    OCF_REQUIRE(bs == 0 || bs >= (long long)C * H * W, OCF_ESHAPE);
    pk.in[t] = xs[t]; pk.in2[t] = nullptr; pk.out[t] = ys[t]; pk.out_bstride[t] = bs;
    vec = vec && ocf_aligned16(xs[t]) && ocf_aligned16(ys[t]) && (bs % 4 == 0);
  }
  const int G = (flags & OCF_NORM_ACROSS_CHANNELS) ? 1 : C;
  const size_t glen = (size_t)(G == 1 ? C : 1) * H * W;
  const long long NGll = (long long)T * B * G;
  OCF_REQUIRE(NGll <= 65535, OCF_EUNSUPPORTED);
  const int NG = (int)NGll;
  vec = vec && (glen % 4 == 0);
  norm_apply_kernel<<<dim3(chunks_for(glen, NG, vec), NG), NT, 0, ocf_cast_stream(stream)>>>(pk, B, G, glen, stats + 6 * (size_t)NG, vec);
  return ocf_launch_status();
}

extern "C" int ocf_normalize_bwd(const float* const* grad_ys, const float* const* xs, float* const* grad_xs, int T, int B, int C, int H,
                                 int W, int flags, const float* stats, float* red, ocf_stream_t stream) {
  OCF_REQUIRE_PTR(grad_ys); OCF_REQUIRE_PTR(xs); OCF_REQUIRE_PTR(grad_xs); OCF_REQUIRE_PTR(stats); OCF_REQUIRE_PTR(red);
  OCF_REQUIRE(T > 0 && T <= MAX_T, OCF_EUNSUPPORTED);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE((flags & ~15) == 0, OCF_EUNSUPPORTED);
  OCF_REQUIRE((reinterpret_cast<uintptr_t>(red) & 7u) == 0, OCF_EALIGN);
  PtrPack pk;
  bool vec = true;
  for (int t = 0; t < T; ++t) {
    OCF_REQUIRE_PTR(grad_ys[t]); OCF_REQUIRE_PTR(xs[t]); OCF_REQUIRE_PTR(grad_xs[t]);
    pk.in[t] = grad_ys[t]; pk.in2[t] = xs[t]; pk.out[t] = grad_xs[t]; pk.out_bstride[t] = 0;
    vec = vec && ocf_aligned16(grad_ys[t]) && ocf_aligned16(xs[t]) && ocf_aligned16(grad_xs[t]);
  }
  const int G = (flags & OCF_NORM_ACROSS_CHANNELS) ? 1 : C;
  const size_t glen = (size_t)(G == 1 ? C : 1) * H * W;
  const long long NGll = (long long)T * B * G;
  OCF_REQUIRE(NGll <= 65535, OCF_EUNSUPPORTED);
  const int NG = (int)NGll;
  cudaStream_t s = ocf_cast_stream(stream);
  vec = vec && (glen % 4 == 0);
  unsigned* ticket = reinterpret_cast<unsigned*>(red + 7 * (size_t)NG);
  cudaError_t e = cudaMemsetAsync(red, 0, sizeof(float) * (8 * (size_t)NG), s);  // fp64 sums and the ticket in one go
  if (e != cudaSuccess) return (int)e;
  const int chunks = chunks_for(glen, NG, vec);
  if (vec) group_sums_kernel<true, true><<<dim3(chunks, NG), NT, 0, s>>>(pk, B, G, glen, red, ticket, stats, NG, flags);
  else group_sums_kernel<true, false><<<dim3(chunks, NG), NT, 0, s>>>(pk, B, G, glen, red, ticket, stats, NG, flags);
  if (int st = ocf_launch_status()) return st;
  norm_bwd_apply_kernel<<<dim3(chunks, NG), NT, 0, s>>>(pk, B, G, glen, stats, red, NG, vec);
  return ocf_launch_status();
}
