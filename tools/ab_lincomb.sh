#!/bin/bash
# A/B of lincomb launch shapes: headline rate + DRAM bytes of the two lincomb kernels (ncu) per library given as arguments
for lib in "$@"; do
  echo "== $lib"
  VRFS_B200_LIB=$PWD/ark_ec_vrfs_b200/$lib python bench.py --steps 5 --warmup 3 --headline-only --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rate', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 2) for k, v in d['roofline']['kernel_ms'].items()})"
  VRFS_B200_LIB=$PWD/ark_ec_vrfs_b200/$lib ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_lincomb -s 8 -c 2 --csv python bench.py --steps 1 --warmup 3 --headline-only --no-cpu-baseline 2>/dev/null | grep -E "k_lincomb" | awk -F'","' '{print $5, $(NF-2), $(NF)}' | cut -c1-200
done
