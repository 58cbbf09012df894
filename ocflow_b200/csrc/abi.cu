// ABI bookkeeping and the host-buffer convenience entry points of include/ocflow_b200.h.
#include "common.cuh"

extern "C" int ocf_abi_version(void) { return OCF_ABI_VERSION; }

extern "C" int ocf_build_sm(void) {
#ifdef OCF_BUILD_SM
  return OCF_BUILD_SM;
#else
  return 100;
#endif
}

extern "C" const char* ocf_error_string(int code) {
  switch (code) {
    case OCF_OK: return "ok";
    case OCF_ENULL: return "OCF_ENULL: a required pointer is NULL";
    case OCF_ESHAPE: return "OCF_ESHAPE: non-positive or inconsistent dimension";
    case OCF_EUNSUPPORTED: return "OCF_EUNSUPPORTED: argument outside what the kernels implement";
    case OCF_EALIGN: return "OCF_EALIGN: misaligned pointer or stride";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown ocflow_b200 error";
}

namespace {

struct DevBuf {
  float* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t n) { return (int)cudaMalloc(&p, n * sizeof(float)); }
};

int h2d(DevBuf& d, const float* h, size_t n) {
  if (int e = d.alloc(n)) return e;
  return (int)cudaMemcpy(d.p, h, n * sizeof(float), cudaMemcpyHostToDevice);
}

}  // namespace

extern "C" int ocf_host_corr_fwd(const float* f1, const float* f2, float* out, int B, int C, int H, int W, int d) {
  OCF_REQUIRE_PTR(f1); OCF_REQUIRE_PTR(f2); OCF_REQUIRE_PTR(out);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  OCF_REQUIRE(d >= 0 && d <= OCF_MAX_DISPLACEMENT, OCF_EUNSUPPORTED);
  const size_t n = (size_t)B * C * H * W, no = (size_t)B * (2 * d + 1) * (2 * d + 1) * H * W;
  DevBuf a, b, o;
  if (int e = h2d(a, f1, n)) return e;
  if (int e = h2d(b, f2, n)) return e;
  if (int e = o.alloc(no)) return e;
  if (int e = ocf_corr_fwd(a.p, b.p, o.p, B, C, H, W, d, 0, 1.0f, nullptr, nullptr, nullptr)) return e;
  return (int)cudaMemcpy(out, o.p, no * sizeof(float), cudaMemcpyDeviceToHost);
}

extern "C" int ocf_host_warp_fwd(const float* img, const float* flow, float* out, int B, int C, int H, int W, int flags) {
  OCF_REQUIRE_PTR(img); OCF_REQUIRE_PTR(flow); OCF_REQUIRE_PTR(out);
  OCF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OCF_ESHAPE);
  const size_t n = (size_t)B * C * H * W, nf = (size_t)B * 2 * H * W;
  DevBuf a, f, o;
  if (int e = h2d(a, img, n)) return e;
  if (int e = h2d(f, flow, nf)) return e;
  if (int e = o.alloc(n)) return e;
  if (int e = ocf_warp_fwd(a.p, f.p, nullptr, o.p, B, C, H, W, flags, 1.0f, nullptr)) return e;
  return (int)cudaMemcpy(out, o.p, n * sizeof(float), cudaMemcpyDeviceToHost);
}

extern "C" int ocf_host_range_map(const float* flow, float* range_out, int B, int H, int W) {
  OCF_REQUIRE_PTR(flow); OCF_REQUIRE_PTR(range_out);
  OCF_REQUIRE(B > 0 && H > 0 && W > 0, OCF_ESHAPE);
  const size_t n = (size_t)B * H * W;
  DevBuf f, o;
  if (int e = h2d(f, flow, 2 * n)) return e;
  if (int e = o.alloc(n)) return e;
  if (int e = ocf_range_map(f.p, o.p, nullptr, B, H, W, nullptr)) return e;
  return (int)cudaMemcpy(range_out, o.p, n * sizeof(float), cudaMemcpyDeviceToHost);
}
