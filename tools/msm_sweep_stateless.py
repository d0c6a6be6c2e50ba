#!/usr/bin/env python3
"""window-size sweep of the stateless 3-column MSM (GLV halves over 2n bases): device ms per (log2 n, c)"""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _scalars import fr_uniform
import ark_ec_vrfs_b200 as vrfs
import oracle_lib as O
lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10, 17)
e = vrfs.Engine(0)
rng = np.random.default_rng(5)
ks = np.zeros((2048, 32), np.uint8); ks[:, :8] = rng.integers(1, 2 ** 62, size=2048, dtype=np.uint64).view(np.uint8).reshape(2048, 8)
base2k = O.g1_mul_gen(ks)
for logn in range(lo, hi + 1):
    n = 1 << logn
    bases = np.tile(base2k, (max(1, n // 2048), 1))[:n]
    sc = fr_uniform(rng, 3 * n)
    ref = None; row = {}
    for c in [0] + list(range(max(7, logn - 4), min(16, logn + 3) + 1)):
        best = None
        for _ in range(3):
            e.enable_kernel_timing(True); out = e.msm_g1(bases, sc, 3, window_bits=c); kt = e.kernel_timings(); e.enable_kernel_timing(False)
            ms = sum(v for _, v in kt)
            if best is None or ms < best[0]: best = (ms, kt)
        if ref is None: ref = out
        assert np.array_equal(ref, out), (logn, c)
        row[c] = round(best[0], 3)
    print("2^%d best:" % logn, min(row, key=row.get), row, flush=True)
