#!/usr/bin/env python3
"""Ring fixed columns + commitments at ring sizes 2^10 / 2^13 / 2^16: one vrfs_ring_commit call from host keys,
Lagrange-basis and monomial SRS (the latter adds the 3-column inverse FFT), kernel and wall-clock ms."""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ark_ec_vrfs_b200 as vrfs
import oracle_lib as O
import vectors as V
e = vrfs.Engine(0)
rng = np.random.default_rng(5)
ks = np.zeros((2048, 32), np.uint8); ks[:, :8] = rng.integers(1, 2 ** 62, size=2048, dtype=np.uint64).view(np.uint8).reshape(2048, 8)
base2k = O.g1_mul_gen(ks)
_, pk0, inp, _ = V.make_keys_inputs(O.BANDERSNATCH, 64)
_, pk_all = e.secret_from_seed(vrfs.BANDERSNATCH, [b"ring-key-%d" % i for i in range(1 << 16)])      # 2^16 distinct public keys
res = {}
for logn in (11, 14, 17):
    n = 1 << logn
    srs = np.tile(base2k, (max(1, n // 2048), 1))[:n]
    keys = pk_all[: n // 2]
    tail = np.tile(inp[:23], (11, 1))
    part = n - 3 - len(tail) - 1
    h = e.msm_g1_prepare(srs)
    row = {}
    for name, lag in (("lagrange", True), ("monomial", False), ("lagrange_delta", None)):
        best = None
        for _ in range(4):
            e.enable_kernel_timing(True); t = time.perf_counter()
            out = h.ring_commit(keys, part, pk0[0], tail, lagrange=lag) if lag is not None else h.ring_commit_delta(keys, pk0[0])
            wall = (time.perf_counter() - t) * 1e3
            kt = e.kernel_timings(); e.enable_kernel_timing(False)
            ms = sum(v for _, v in kt)
            if best is None or ms < best[0]: best = (ms, wall, kt)
        row[name] = {"kernels_ms": round(best[0], 3), "wall_ms": round(best[1], 3), "per_kernel": {a: round(b, 3) for a, b in best[2]}}
        print("ring size 2^%d (domain 2^%d) %-9s kernels %.3f ms, wall %.3f ms" % (logn - 1, logn, name, best[0], best[1]), {a: round(b, 2) for a, b in best[2]}, flush=True)
    cols = e.ring_fixed_columns(n, part, keys, pk0[0], tail).reshape(-1, 32)
    t = time.perf_counter(); e.fr_fft(cols, 3, inverse=True); row["ifft_3col_wall_ms"] = round((time.perf_counter() - t) * 1e3, 3)
    # SRS / commitment points on the wire: compress, decompress with and without the subgroup test (kernel ms)
    enc = e.g1_compress(srs)
    for name, f in (("g1_compress", lambda: e.g1_compress(srs)), ("g1_decompress_checked", lambda: e.g1_decompress(enc, True)), ("g1_decompress_unchecked", lambda: e.g1_decompress(enc, False))):
        e.enable_kernel_timing(True); f(); row[name + "_kernel_ms"] = round(sum(v for _, v in e.kernel_timings()), 3); e.enable_kernel_timing(False)
    print("   G1 wire, %d points: compress %.3f ms, decompress %.3f ms (with subgroup test) / %.3f ms (without)" % (n, row["g1_compress_kernel_ms"], row["g1_decompress_checked_kernel_ms"], row["g1_decompress_unchecked_kernel_ms"]), flush=True)
    h.release()
    res["2^%d" % logn] = row
json.dump(res, open(os.path.join(ROOT, "gpurun_out", os.environ.get("RING_BENCH_OUT", "ring_bench.json")), "w"), indent=1)
