#!/usr/bin/env python3
"""One prepared 3-column MSM at N = 2^17 (ring size 2^16) for ncu captures."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _scalars import fr_uniform
import ark_ec_vrfs_b200 as vrfs
import oracle_lib as O
n = 1 << int(os.environ.get("MSM_LOGN", "17"))
rng = np.random.default_rng(5)
ks = np.zeros((2048, 32), np.uint8); ks[:, :8] = rng.integers(1, 2 ** 62, size=2048, dtype=np.uint64).view(np.uint8).reshape(2048, 8)
bases = np.tile(O.g1_mul_gen(ks), (n // 2048, 1))
sc = fr_uniform(rng, 3 * n)
e = vrfs.Engine(0)
h = e.msm_g1_prepare(bases)
for _ in range(2):
    out = h.msm(sc, 3)
print(out[0, :8])
