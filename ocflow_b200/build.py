"""Builds ocflow_b200/libocflow_b200.so (C ABI, include/ocflow_b200.h) with nvcc for sm_100a only.

    python -m ocflow_b200.build [--force] [--verbose]

The library is built IN-TREE (git-ignored, travels to the GPU box with the snapshot).  nvcc
cross-compiles without a GPU.  There is deliberately no other architecture in the fat binary.
Every .cu file is compiled to its own object (in parallel, re-compiled only when it or a header changed),
then linked; `--force` rebuilds everything from scratch.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libocflow_b200.so")
STAMP = os.path.join(PKG, ".libocflow_b200.stamp")
SOURCES = ["corr.cu", "corr_tc.cu", "level.cu", "warp.cu", "resample.cu", "conv_glue.cu", "loss.cu", "normalize.cu", "ssim.cu", "census.cu", "metrics.cu",
           "pack.cu", "abi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-DOCF_BUILD_SM=100",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _headers_digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    files.append(os.path.join(PKG, "..", "include", "ocflow_b200.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _file_digest(src, hdr):
    h = hashlib.sha256(hdr.encode())
    with open(os.path.join(CSRC, src), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def _digest():
    hdr = _headers_digest()
    return hashlib.sha256("".join(_file_digest(s, hdr) for s in sources()).encode()).hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != _digest()


def _compile_one(src, hdr, force, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    stamp = obj + ".stamp"
    want = _file_digest(src, hdr)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return obj, ""
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed on %s (exit %d):\n%s%s" % (src, proc.returncode, proc.stdout, proc.stderr))
    with open(stamp, "w") as fh:
        fh.write(want)
    return obj, proc.stdout + proc.stderr


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers_digest()
    srcs = sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda s: _compile_one(s, hdr, force, verbose), srcs))
    if verbose:
        for (_, log), s in zip(results, srcs):
            if log:
                sys.stderr.write("==== %s\n%s" % (s, log))
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + [o for o, _ in results]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed linking libocflow_b200.so (exit %d)" % proc.returncode)
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
