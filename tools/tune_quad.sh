#!/bin/bash
# Developer tuning run (GPU box): quad gather/scatter paths on/off, CTA targets, fused-loss occupancy.
out=gpurun_out/r2c_tune.txt
: > $out
for cfg in "OCF_WARP_QUAD=0" "OCF_WARP_QUAD=1 OCF_QUAD_CTAS=1" "OCF_WARP_QUAD=1 OCF_QUAD_CTAS=2" "OCF_WARP_QUAD=1 OCF_QUAD_CTAS=4" "OCF_WARP_QUAD=2 OCF_QUAD_CTAS=2" "OCF_WARP_QUAD=2 OCF_QUAD_CTAS=4" "OCF_WARP_QUAD=4" "OCF_OPF_MINB=3"; do
  echo "==== $cfg" >> $out
  env $cfg python bench.py --kernels-only --only warp_,range_map,occ_photo 2>/dev/null | grep -v "L6\|L5" | cut -c1-75 >> $out
done
