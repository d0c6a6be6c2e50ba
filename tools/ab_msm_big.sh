for lib in "$@"; do
  echo "== $lib"
  VRFS_B200_LIB=$PWD/ark_ec_vrfs_b200/$lib MSM_ONLY=1 python tools/bench_all.py --msm-max-logn 17 --out gpurun_out/x.json 2>&1 | grep -E "2\^1[567] " | sed "s/stateless \([0-9.]*\) ms (device \([0-9.]*\)).*prepared \([0-9.]*\) ms (device \([0-9.]*\)).*msm_accumulate.: \([0-9.]*\).*/stateless dev \2  prepared dev \4  acc \5/"
done
