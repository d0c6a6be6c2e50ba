#!/bin/bash
# A/B: run bench.py (short) against each libvrfs_*.so variant given as arguments (names inside ark_ec_vrfs_b200/)
for lib in "$@"; do
  echo "== $lib"
  VRFS_B200_LIB=$PWD/ark_ec_vrfs_b200/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3), {k:round(v,2) for k,v in d['roofline']['kernel_ms'].items()}, 'frac',round(d['roofline']['frac'],3), d['clocks'])"
done
