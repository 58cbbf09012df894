#!/bin/bash
# builds tools/bin/lib_<name>.so with extra -D flags: tools/build_variant.sh name -DOCF_WB_CB=4 ...   (load it with OCFLOW_B200_LIB)
# The source list comes from ocflow_b200.build (every entry point must be present: _lib.load() binds all of them); nvcc's
# exit status is propagated.
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p tools/bin
srcs=$(python -c "from ocflow_b200 import build as b; print(' '.join('ocflow_b200/csrc/' + s for s in b.sources()))")
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DOCF_BUILD_SM=100 -Xcompiler -fPIC -shared "$@" -o tools/bin/lib_$name.so $srcs
