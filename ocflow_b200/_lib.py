"""ctypes binding of libocflow_b200.so (the C ABI declared in include/ocflow_b200.h).

The library is the ONLY compute path of this package: there is no CPU fallback and no other backend.
If it is missing the import of any op fails loudly (RuntimeError) instead of degrading.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# OCFLOW_B200_LIB: developer override used by the tuning scripts in tools/ to load an alternative build of the SAME library
LIB_PATH = os.environ.get("OCFLOW_B200_LIB") or os.path.join(_PKG, "libocflow_b200.so")

c_f = ctypes.c_void_p  # device (or host) float* / double* passed as raw addresses
c_i = ctypes.c_int
c_ll = ctypes.c_longlong
c_fl = ctypes.c_float
c_s = ctypes.c_void_p  # cudaStream_t

# name -> argtypes ; every entry returns int.  Must list every symbol of include/ocflow_b200.h
# (tests/test_abi.py cross-checks this table against the header).
SIGNATURES = {
    "ocf_corr_fwd": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_ll, c_fl, c_f, c_f, c_s],
    "ocf_corr_bwd": [c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_ll, c_ll, c_fl, c_f, c_s],
    "ocf_normalize_fwd": [c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_s],
    "ocf_normalize_stats": [c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_s],
    "ocf_normalize_apply": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_s],
    "ocf_corr_fwd_strided": [c_f, c_ll, c_f, c_f, c_ll, c_f, c_i, c_i, c_i, c_i, c_fl, c_s],
    "ocf_level_corr_fwd": [c_f, c_f, c_f, c_f, c_ll, c_f, c_ll, c_f, c_f, c_i, c_i, c_i, c_i, c_fl, c_s],
    "ocf_level_corr_bwd": [c_f, c_ll, c_f, c_f, c_ll, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_fl, c_s],
    "ocf_normalize_bwd": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_s],
    "ocf_resize_bilinear_fwd": [c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_fl, c_s],
    "ocf_resize_bilinear_bwd": [c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_fl, c_s],
    "ocf_bias_lrelu_fwd": [c_f, c_f, c_f, c_i, c_i, c_ll, c_fl, c_s],
    "ocf_bias_lrelu_bwd": [c_f, c_f, c_f, c_f, c_i, c_i, c_ll, c_fl, c_s],
    "ocf_warp_fwd": [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_fl, c_s],
    "ocf_warp_bwd": [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_fl, c_s],
    "ocf_range_map": [c_f, c_f, c_f, c_i, c_i, c_i, c_s],
    "ocf_flow_to_warp": [c_f, c_f, c_i, c_i, c_i, c_s],
    "ocf_robust_l1_fwd": [c_f, c_f, c_ll, c_fl, c_s],
    "ocf_robust_l1_bwd": [c_f, c_f, c_f, c_ll, c_fl, c_s],
    "ocf_photometric_fwd": [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_fl, c_s],
    "ocf_photometric_bwd": [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_fl, c_s],
    "ocf_smooth_fwd": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_fl, c_fl, c_s],
    "ocf_smooth_bwd": [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_fl, c_fl, c_s],
    "ocf_gradient": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_s],
    "ocf_occ_photo_fused": [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_fl, c_s],
    "ocf_pair_loss": [c_f, c_f, c_f, c_f, c_ll, c_i, c_s],
    "ocf_ssim_fwd": [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_s],
    "ocf_ssim_bwd": [c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_s],
    "ocf_census_fwd": [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_s],
    "ocf_census_bwd": [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_s],
    "ocf_flow_metrics": [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_s],
    "ocf_pack_occ": [c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_s],
    "ocf_pack_pairs": [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_s],
    "ocf_host_corr_fwd": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i],
    "ocf_host_warp_fwd": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i],
    "ocf_host_range_map": [c_f, c_f, c_i, c_i, c_i],
}
NOARG = {"ocf_abi_version": c_i, "ocf_build_sm": c_i}

_lib = None


def load():
    """dlopen the library (once) and attach prototypes.  Raises RuntimeError when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "ocflow_b200: %s is missing -- build it with `python -m ocflow_b200.build` "
            "(or __graft_entry__.build()); there is no fallback path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_i
    for name, res in NOARG.items():
        fn = getattr(lib, name)
        fn.argtypes = []
        fn.restype = res
    lib.ocf_error_string.argtypes = [c_i]
    lib.ocf_error_string.restype = ctypes.c_char_p
    _lib = lib
    return lib


def error_string(code):
    return load().ocf_error_string(int(code)).decode()


# launch counter: every successful C-ABI call that enqueues kernels bumps it (bench.py's gpu_launches)
launch_count = 0
KERNELS_PER_CALL = {
    "ocf_corr_fwd": 1, "ocf_corr_bwd": 1, "ocf_normalize_stats": 1, "ocf_normalize_apply": 1, "ocf_corr_fwd_strided": 1, "ocf_level_corr_fwd": 1, "ocf_level_corr_bwd": 1, "ocf_normalize_fwd": 2, "ocf_normalize_bwd": 2, "ocf_warp_fwd": 1, "ocf_bias_lrelu_fwd": 1, "ocf_bias_lrelu_bwd": 1, "ocf_resize_bilinear_fwd": 1, "ocf_resize_bilinear_bwd": 1,
    "ocf_warp_bwd": 1, "ocf_range_map": 1, "ocf_flow_to_warp": 1, "ocf_robust_l1_fwd": 1, "ocf_robust_l1_bwd": 1,
    "ocf_photometric_fwd": 1, "ocf_photometric_bwd": 1, "ocf_smooth_fwd": 1, "ocf_smooth_bwd": 1, "ocf_gradient": 1,
    "ocf_occ_photo_fused": 1, "ocf_pair_loss": 1, "ocf_ssim_fwd": 1, "ocf_ssim_bwd": 1, "ocf_census_fwd": 1, "ocf_census_bwd": 1, "ocf_flow_metrics": 1, "ocf_pack_pairs": 1, "ocf_pack_occ": 1,
}


# optional live timer (bench.py): when set to a list, every call is bracketed by CUDA events recorded on the stream the
# kernels are enqueued on (torch's current stream == the stream argument); entries are (name, int_args, start, end).
live_timer = None


def call(name, *args):
    """Invoke an entry point; raise RuntimeError with the library's message on a non-zero status."""
    global launch_count
    if live_timer is not None:
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        code = getattr(load(), name)(*args)
        e1.record()
        live_timer.append((name, tuple(a for a in args if isinstance(a, int)), e0, e1))
    else:
        code = getattr(load(), name)(*args)
    if code != 0:
        raise RuntimeError("ocflow_b200.%s failed: %s (code %d)" % (name, error_string(code), code))
    launch_count += KERNELS_PER_CALL.get(name, 0)
    return 0
