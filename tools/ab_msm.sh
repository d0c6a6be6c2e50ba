#!/bin/bash
# A/B of MSM variants: prepared 3-column MSM timings per library given as arguments
for lib in "$@"; do
  echo "== $lib"
  VRFS_B200_LIB=$PWD/ark_ec_vrfs_b200/$lib MSM_ONLY=1 python tools/bench_all.py --msm-max-logn 17 --out /tmp/ab_msm.json 2>&1 | grep -E "2\^1[1457] " | sed 's/{.*msm_accumulate.: \([0-9.]*\).*/acc \1/'
done
