import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import ark_ec_vrfs_b200 as vrfs
n=1<<20
with vrfs.Engine(0) as e:
    seeds=[b"w"+i.to_bytes(8,"little") for i in range(n)]
    sk,pk=e.secret_from_seed(0,seeds)
    enc=e.point_encode(0,pk)
    e.enable_kernel_timing(True)
    p1=e.point_decode(0,enc); k1=dict(e.kernel_timings())
    p2=e.point_decode_checked(0,enc); k2=dict(e.kernel_timings())
    ok=e.subgroup_check(0,pk); k3=dict(e.kernel_timings())
    print("decode", k1, "decode_checked", k2, "subgroup_check", k3)
