#!/usr/bin/env python3
"""Wall-clock latency of vrfs_ietf_verify_batch (host buffers) for small batches: what a per-item caller of the drop-in sees."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ark_ec_vrfs_b200 as vrfs
import oracle_lib as O, vectors as V
e = vrfs.Engine(0)
w = V.make_ietf_proofs(0, 4096, "empty")
for n in (1, 16, 256, 4096, 65536):
    reps = max(1, n // 4096)
    a = {k: np.ascontiguousarray(np.tile(w[k], (reps, 1))[:n]) for k in ("pk", "inp", "out", "c", "s")}
    for _ in range(3):
        got = e.ietf_verify(0, a["pk"], a["inp"], a["out"], a["c"], a["s"], None)
    ts = []
    for _ in range(10):
        t = time.perf_counter(); got = e.ietf_verify(0, a["pk"], a["inp"], a["out"], a["c"], a["s"], None); ts.append(time.perf_counter() - t)
    assert np.array_equal(got, np.tile(w["expect"], reps)[:n])
    print("n = %6d: %.3f ms per call (median of 10), %.1f us per proof" % (n, sorted(ts)[5] * 1e3, sorted(ts)[5] / n * 1e6), flush=True)
t = time.perf_counter(); O.ietf_verify(0, w["pk"][:256], w["inp"][:256], w["out"][:256], w["c"][:256], w["s"][:256], None, nthreads=1); dt = time.perf_counter() - t
print("CPU oracle, 1 thread: %.1f us per proof" % (dt / 256 * 1e6))
