#!/bin/bash
# developer aid: sweep the load-batch knobs of the warp / range-map kernels (run on the GPU box)
for cfg in "OCF_WARP_CB=88 OCF_RANGE_PPT=4" "OCF_WARP_CB=44 OCF_RANGE_PPT=1" "OCF_WARP_CB=48 OCF_RANGE_PPT=2" "OCF_WARP_CB=84 OCF_RANGE_PPT=8"; do
  echo "==== $cfg"
  env $cfg python bench.py --kernels-only --only warp_,range_map 2>/dev/null | grep -v "^{" 
done
