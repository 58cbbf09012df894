"""Builds ocflow_b200/libocflow_b200.so (C ABI, include/ocflow_b200.h) with nvcc for sm_100a only.

    python -m ocflow_b200.build [--force] [--verbose]

The library is built IN-TREE (git-ignored, travels to the GPU box with the snapshot).  nvcc
cross-compiles without a GPU.  There is deliberately no other architecture in the fat binary.
"""
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libocflow_b200.so")
STAMP = os.path.join(PKG, ".libocflow_b200.stamp")
SOURCES = ["corr.cu", "warp.cu", "loss.cu", "normalize.cu", "ssim.cu", "census.cu", "metrics.cu", "pack.cu", "abi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-DOCF_BUILD_SM=100",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(PKG, "..", "include", "ocflow_b200.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != _digest()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    flags = list(NVCC_FLAGS)
    cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libocflow_b200.so (exit %d)" % proc.returncode)
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
