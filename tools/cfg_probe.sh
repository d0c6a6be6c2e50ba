for lib in libvrfs_b200.so libvrfs_t128_b3.so libvrfs_t128_b5.so libvrfs_t256_b2.so; do
  echo "== $lib"
  VRFS_B200_LIB=$PWD/ark_ec_vrfs_b200/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3),'M/s', {k:round(v,2) for k,v in d['roofline']['kernel_ms'].items()}, 'frac',round(d['roofline']['frac'],3), d['roofline']['peak_probe_tmac32'])"
done
