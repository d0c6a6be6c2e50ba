import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np
import ark_ec_vrfs_b200 as vrfs
eng = vrfs.Engine(0)
n = 1 << 20
sk256, pk256 = eng.secret_from_seed(0, [b"bench-wire-%d" % i for i in range(256)])
sk = np.tile(sk256, (n // 256, 1)); pk_enc = np.tile(eng.point_encode(0, pk256), (n // 256, 1))
datas = (np.arange(n, dtype=np.uint64).view(np.uint8).copy(), np.arange(n + 1, dtype=np.uint64) * 8)
sig, ok = eng.ietf_sign_wire(0, sk, datas); eng.ietf_verify_wire(0, pk_enc, datas, sig)   # warm-up: buffers, module load
eng.enable_kernel_timing(True)
t0=time.perf_counter(); sig, ok = eng.ietf_sign_wire(0, sk, datas); t1=time.perf_counter()
print("sign", n/(t1-t0)/1e6, "M/s", eng.kernel_timings())
for _ in range(2):
    t0=time.perf_counter(); okv, beta = eng.ietf_verify_wire(0, pk_enc, datas, sig); t1=time.perf_counter()
    print("verify_wire", n/(t1-t0)/1e6, "M/s", eng.kernel_timings())
