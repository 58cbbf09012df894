#!/bin/bash
# builds tools/bin/lib_<name>.so with extra -D flags: tools/build_variant.sh name -DOCF_WB_CB=4 ...
name=$1; shift
mkdir -p tools/bin
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DOCF_BUILD_SM=100 -Xcompiler -fPIC -shared "$@" \
  -o tools/bin/lib_$name.so ocflow_b200/csrc/{corr,warp,loss,normalize,ssim,census,metrics,abi}.cu 2>&1 | grep -E "error" 
