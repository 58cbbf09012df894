#!/bin/bash
# times corr fwd/bwd at every pyramid level for each channel-split factor (developer aid)
for ks in 1 2 4 8; do
  echo "== ksplit $ks"
  OCF_KSPLIT_FWD=$ks OCF_KSPLIT_BWD=$ks python bench.py --kernels-only 2>&1 | grep corr_ | cut -c1-60
done
