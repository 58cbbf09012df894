"""CPU: the C-ABI library loads, exports every symbol include/ocflow_b200.h declares, and validates arguments
before touching the GPU (no compute call is made here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "ocflow_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = re.findall(r"\b(?:int|const char\*)\s+(ocf_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S)
    return {name: [a.strip() for a in args.split(",")] if args.strip() != "void" else [] for name, args in decls}


def test_header_declares_and_library_exports_every_symbol():
    from ocflow_b200 import _lib

    decl = _declared()
    assert len(decl) >= 20
    lib = _lib.load()
    for name in decl:
        assert hasattr(lib, name), "libocflow_b200.so does not export %s" % name
    assert lib.ocf_abi_version() == 3
    assert lib.ocf_build_sm() == 100
    assert b"OCF_ENULL" in lib.ocf_error_string(-1)


def test_ctypes_table_matches_header_arity():
    from ocflow_b200 import _lib

    decl = _declared()
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in decl, name
        assert len(decl[name]) == len(argtypes), "%s: header has %d args, ctypes table %d" % (name, len(decl[name]), len(argtypes))
    known = set(_lib.SIGNATURES) | set(_lib.NOARG) | {"ocf_error_string"}
    assert set(decl) == known, set(decl) ^ known


def test_argument_validation_happens_before_any_launch():
    from ocflow_b200 import _lib

    lib = _lib.load()
    one = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    assert lib.ocf_corr_fwd(None, one, one, 1, 1, 1, 1, 4, 0, 1.0, None, None, None) == -1
    assert lib.ocf_corr_fwd(one, one, one, 0, 1, 1, 1, 4, 0, 1.0, None, None, None) == -2
    assert lib.ocf_corr_fwd(one, one, one, 1, 1, 1, 1, 17, 0, 1.0, None, None, None) == -3
    assert lib.ocf_corr_fwd(one, one, one, 1, 1, 1, 1, 2, 0, 1.0, None, one, None) == -3   # sign mask only from the d = 4 kernels
    assert lib.ocf_corr_fwd(one, one, one, 1, 1, 2, 2, 4, 5, 1.0, None, None, None) == -2  # out_bstride too small
    assert lib.ocf_corr_bwd(one, None, one, one, None, None, 1, 1, 1, 1, 4, 0, 0, 1.0, None, None) == -1
    assert lib.ocf_warp_fwd(one, one, None, one, 1, 1, 1, 1, 8, 1.0, None) == -3
    assert lib.ocf_warp_bwd(one, one, one, None, None, None, None, 1, 1, 1, 1, 0, 1.0, None) == -1
    assert lib.ocf_range_map(one, None, None, 1, 1, 1, None) == -1
    assert lib.ocf_smooth_fwd(one, one, one, 1, 3, 2, 4, 4, 3, 100.0, 0.001, None) == -3
    assert lib.ocf_pair_loss(one, one, one, None, 4, 9, None) == -3
    assert lib.ocf_host_corr_fwd(None, None, None, 1, 1, 1, 1, 4) == -1


def test_every_entry_point_rejects_null_pointers_and_empty_shapes_without_a_device():
    """The whole ABI, not a sample: all-NULL arguments give OCF_ENULL, dummy (never dereferenced) pointers with zero sizes give
    OCF_ESHAPE / OCF_EUNSUPPORTED -- each before anything is enqueued, so this runs where there is no GPU."""
    from ocflow_b200 import _lib

    lib = _lib.load()
    dummy = ctypes.c_void_p(4096)
    for name, argtypes in _lib.SIGNATURES.items():
        def args(ptr):
            return [ptr if t is ctypes.c_void_p else (0.0 if t in (ctypes.c_float, ctypes.c_double) else 0) for t in argtypes]
        fn = getattr(lib, name)
        assert fn(*args(None)) == -1, name
        assert fn(*args(dummy)) in (-2, -3), name


def test_missing_library_fails_loudly(monkeypatch):
    from ocflow_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libocflow_b200.so")
    with pytest.raises(RuntimeError, match="no fallback"):
        _lib.load()


def test_header_is_plain_c_and_links_from_a_c_program(tmp_path):
    """The boundary is a C ABI: the header must compile as C (not only C++), and a C program must link against the shared
    library and get the documented error codes back -- without touching a GPU."""
    import shutil
    import subprocess

    from ocflow_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi_smoke.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "ocflow_b200.h"\n'
        "int main(void) {\n"
        "  float x[4] = {0};\n"
        "  if (ocf_abi_version() != OCF_ABI_VERSION) return 1;\n"
        "  if (ocf_build_sm() != 100) return 2;\n"
        "  if (ocf_corr_fwd(0, x, x, 1, 1, 1, 1, 4, 0, 1.0f, 0, 0, 0) != OCF_ENULL) return 3;\n"
        "  if (ocf_corr_fwd(x, x, x, 1, 1, 1, 1, 99, 0, 1.0f, 0, 0, 0) != OCF_EUNSUPPORTED) return 4;\n"
        "  if (ocf_warp_fwd(x, x, 0, x, 1, 1, 0, 1, 0, 1.0f, 0) != OCF_ESHAPE) return 5;\n"
        "  if (ocf_flow_metrics(x, x, x, 0, 1, 1, 1, 0, 0) != OCF_ENULL) return 6;\n"
        "  if (ocf_pack_pairs(0, 0, 0, 0, 0, 1, 4, 4, 4, 4, 0, 0, 0) != OCF_ENULL) return 7;\n"
        '  if (strstr(ocf_error_string(OCF_ESHAPE), "OCF_ESHAPE") == 0) return 8;\n'
        '  printf("abi ok\\n");\n  return 0;\n}\n')
    exe = tmp_path / "abi_smoke"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cc = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                         "-L", libdir, "-locflow_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and "abi ok" in run.stdout, (run.returncode, run.stdout, run.stderr)
