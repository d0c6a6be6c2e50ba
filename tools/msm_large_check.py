#!/usr/bin/env python3
"""Prepared MSM beyond the ring sizes (2^18 and 2^20 points, one column) with a known answer: the bases and scalars repeat with
period 2048, so the sum is (n / 2048) times a 2048-point MSM that the oracle computes - also an extreme-skew case (every digit
pattern repeats n / 2048 times)."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import ark_ec_vrfs_b200 as vrfs
import oracle_lib as O
from _scalars import fr_uniform
e = vrfs.Engine(0)
rng = np.random.default_rng(5)
ks = np.zeros((2048, 32), np.uint8); ks[:, :8] = rng.integers(1, 2 ** 62, size=2048, dtype=np.uint64).view(np.uint8).reshape(2048, 8)
base2k = O.g1_mul_gen(ks)
for logn in (18, 20):
    n = 1 << logn
    bases = np.tile(base2k, (n // 2048, 1))
    # columns built so that the answer is known: scalar j is s on rows j = i (mod 2048) pattern -> sum = (n/2048) * sum_i s_i P_i
    s2k = fr_uniform(rng, 2048)
    sc = np.tile(s2k, (n // 2048, 1))
    h = e.msm_g1_prepare(bases)
    t = time.perf_counter(); out = h.msm(sc, 1); dt = time.perf_counter() - t
    h.release()
    # expected: (n/2048) * MSM_2048(s2k) computed at the small size: scale the scalars by n/2048 mod r
    R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    scaled = np.frombuffer(b"".join(((int.from_bytes(r.tobytes(), "little") * (n // 2048)) % R).to_bytes(32, "little") for r in s2k), np.uint8).reshape(-1, 32)
    exp = O.msm_g1(base2k, scaled, 1)
    print("2^%d: %.2f ms, correct = %s" % (logn, dt * 1e3, np.array_equal(out, exp)), flush=True)
