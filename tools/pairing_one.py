"""one product of two pairings through the lane kernel (profiling target)"""
import sys, random
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import ark_ec_vrfs_b200 as vrfs
from oracle import pairing_ref as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rnd = random.Random(1)
a, b = rnd.randrange(1, P.R), rnd.randrange(1, P.R)
g1 = P.g1_to_bytes(P.g1_mul(a, P.G1_GEN)) + P.g1_to_bytes(P.g1_mul(a * b, P.G1_GEN))
g2 = P.g2_to_bytes(P.g2_mul(b, P.G2_GEN)) + P.g2_to_bytes(P.G2_GEN)
u8 = lambda x: np.frombuffer(bytes(x), np.uint8)
with vrfs.Engine(0) as eng:
    for _ in range(2):
        ok = eng.pairing_products(u8(g1 * n), u8(g2 * n), 2, negate_masks=[2] * n)
    assert ok.tolist() == [1] * n
