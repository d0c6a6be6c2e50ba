#!/usr/bin/env python3
"""Prepared 3-column MSM with the column shapes of a ring commitment: two columns of random 255-bit scalars (x, y of the keys)
and one 0/1 selector column (ones on the first half of the domain) - the skewed case (one bucket holds n/2 entries)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _scalars import fr_uniform
import ark_ec_vrfs_b200 as vrfs
import oracle_lib as O
e = vrfs.Engine(0)
rng = np.random.default_rng(5)
ks = np.zeros((2048, 32), np.uint8); ks[:, :8] = rng.integers(1, 2 ** 62, size=2048, dtype=np.uint64).view(np.uint8).reshape(2048, 8)
base2k = O.g1_mul_gen(ks)
for logn in (11, 14, 17):
    n = 1 << logn
    bases = np.tile(base2k, (max(1, n // 2048), 1))[:n]
    sc = fr_uniform(rng, 3 * n)
    sel = sc.copy(); sel[2 * n:] = 0; sel[2 * n:2 * n + n // 2, 0] = 1
    h = e.msm_g1_prepare(bases)
    for name, s in (("random x3", sc), ("x, y, selector", sel)):
        best = None
        for _ in range(4):
            e.enable_kernel_timing(True); out = h.msm(s, 3); kt = e.kernel_timings(); e.enable_kernel_timing(False)
            ms = sum(v for _, v in kt)
            if best is None or ms < best[0]: best = (ms, kt)
        print("2^%d %-15s %.3f ms" % (logn, name, best[0]), {a: round(b, 2) for a, b in best[1]}, flush=True)
    if logn <= 11:
        assert np.array_equal(out, O.msm_g1(bases, sel, 3))
    h.release()
