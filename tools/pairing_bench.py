"""GPU timing of the pairing product paths: one warp per product (lane programs) vs one thread per product, and the KZG check."""
import json, sys, time, random
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import ark_ec_vrfs_b200 as vrfs
from oracle import pairing_ref as P

def u8(b): return np.frombuffer(bytes(b), np.uint8)
rnd = random.Random(1)
a, b = rnd.randrange(1, P.R), rnd.randrange(1, P.R)
g1 = P.g1_to_bytes(P.g1_mul(a, P.G1_GEN)) + P.g1_to_bytes(P.g1_mul(a * b, P.G1_GEN))
g2 = P.g2_to_bytes(P.g2_mul(b, P.G2_GEN)) + P.g2_to_bytes(P.G2_GEN)
res = {}
with vrfs.Engine(0) as eng:
    eng.enable_kernel_timing(True)
    for n in (1, 148, 592, 1184, 4736, 8192):
        G1, G2 = u8(g1 * n), u8(g2 * n)
        best = 1e9
        for _ in range(3):
            t = time.perf_counter(); ok = eng.pairing_products(G1, G2, 2, negate_masks=[2] * n); dt = time.perf_counter() - t
            best = min(best, dt)
        assert ok.tolist() == [1] * n
        res[n] = {"wall_ms": best * 1e3, "kernel_ms": eng.kernel_timings(), "products_per_s": n / best}
        print(n, res[n], flush=True)
json.dump(res, open("gpurun_out/pairing_bench.json", "w"), indent=1)
