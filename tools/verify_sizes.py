#!/usr/bin/env python3
"""Bandersnatch IETF verify through the host-buffer call at batch sizes 2^15 .. 2^20 (one GPU): M verifies/s per size.
The sizes between one and a few resident waves of the lincomb grid are what a multi-GPU caller hands each GPU."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ark_ec_vrfs_b200 as vrfs
n = 1 << 20
with vrfs.Engine(0) as e:
    seeds = [b"vs-sk" + i.to_bytes(8, "little") for i in range(n)]
    alphas = [i.to_bytes(8, "little") + bytes(24) for i in range(n)]
    sk, pk = e.secret_from_seed(0, seeds)
    inp, ok = e.data_to_point(0, alphas)
    out = e.output(0, sk, inp)
    c, s = e.ietf_prove(0, sk, inp, out)
    for lg in (15, 16, 17, 18, 19, 20):
        m = 1 << lg
        best = 1e9
        for _ in range(5):
            t = time.perf_counter(); okv = e.ietf_verify(0, pk[:m], inp[:m], out[:m], c[:m], s[:m]); best = min(best, time.perf_counter() - t)
        assert okv.all()
        print("2^%d: %.3f ms  %.2f M verifies/s" % (lg, best * 1e3, m / best / 1e6), flush=True)
