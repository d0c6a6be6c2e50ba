#!/usr/bin/env python3
"""(window bits, threads per bucket) sweep of the prepared 3-column MSM (tuning aid for msm_plan): device ms per (log2 n, c, tpb).
  python tools/msm_sweep2.py [lo hi]"""
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
res = {}
for logn in range(lo, hi + 1):
    n = 1 << logn
    bases = np.tile(base2k, (max(1, n // 2048), 1))[:n]
    sc = fr_uniform(rng, 3 * n)
    ref = None; row = {}
    cs = [0] + [c for c in (8, 9, 10, 11, 13, 15, 16) if logn - 3 <= c <= logn + 4]
    for c in cs:
        for tpb in ([0] if c == 0 else [1, 2, 4, 8, 16, 32]):
            h = e.msm_g1_prepare(bases, window_bits=c, threads_per_bucket=tpb)
            best = None
            for _ in range(3):
                e.enable_kernel_timing(True); out = h.msm(sc, 3); kt = e.kernel_timings(); e.enable_kernel_timing(False)
                ms = sum(v for _, v in kt)
                if best is None or ms < best[0]: best = (ms, kt)
            h.release()
            if ref is None: ref = out
            assert np.array_equal(ref, out), (logn, c, tpb)
            row["%d/%d" % (c, tpb)] = round(best[0], 3)
            print("2^%d c=%d tpb=%d: %.3f ms" % (logn, c, tpb, best[0]), {a: round(b, 3) for a, b in best[1] if b > 0.02}, flush=True)
    res[logn] = row
    print("2^%d best:" % logn, min(row, key=row.get), sorted(row.items(), key=lambda kv: kv[1])[:5], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "msm_sweep2.json"), "w"), indent=1)
