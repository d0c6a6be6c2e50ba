import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, time
import ark_ec_vrfs_b200 as vrfs
e = vrfs.Engine(0)
n = 1 << 20
seeds = [b"k%d" % i for i in range(256)]
sk256, pk256 = e.secret_from_seed(0, seeds)
sk = np.tile(sk256, (n // 256, 1))
datas = (np.arange(n, dtype=np.uint64).view(np.uint8).copy(), np.arange(n + 1, dtype=np.uint64) * 8)
inp, ok = e.data_to_point(0, datas)
out = e.output(0, sk, inp)
pr, bl = e.pedersen_prove(0, sk, inp, out)
e.enable_kernel_timing(True)
pr, bl = e.pedersen_prove(0, sk, inp, out); print("prove", e.kernel_timings())
okp = e.pedersen_verify(0, inp, out, pr); print("verify", e.kernel_timings(), okp.all())
c, s = e.ietf_prove(0, sk, inp, out); print("ietf prove", e.kernel_timings())
