#!/usr/bin/env python3
"""secp256r1 IETF prove / verify of 2^20 items (device-resident kernels: sum of the call's kernel times), one GPU"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ark_ec_vrfs_b200 as vrfs
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
with vrfs.Engine(0) as e:
    for suite, name in ((2, "secp256r1"), (1, "ed25519"), (0, "bandersnatch")):
        seeds = [b"pb-sk" + i.to_bytes(8, "little") for i in range(n)]
        alphas = [i.to_bytes(8, "little") + bytes(24) for i in range(n)]
        sk, pk = e.secret_from_seed(suite, seeds)
        inp, ok = e.data_to_point(suite, alphas)
        out = e.output(suite, sk, inp)
        sk, pk, inp, out = (vrfs.host_copy(x) for x in (sk, pk, inp, out))   # page-locked: the copies of a piece-wise call run beside its kernels
        e.ietf_prove(suite, sk, inp, out)                                    # warm-up (fixed-base tables, buffer growth)
        e.enable_kernel_timing(True)
        def summed():          # a piece-wise host call lists every kernel once per piece
            d = {}
            for k, ms in e.kernel_timings():
                d[k] = d.get(k, 0.0) + ms
            return d
        c, s = e.ietf_prove(suite, sk, inp, out); kp = summed()
        okv = e.ietf_verify(suite, pk, inp, out, c, s); kv = summed()
        e.enable_kernel_timing(False)
        assert okv.all()
        print("%-12s prove %.2f M/s %s   verify %.2f M/s %s" % (name, n / sum(kp.values()) / 1e3, {k: round(v, 1) for k, v in kp.items()},
                                                                n / sum(kv.values()) / 1e3, {k: round(v, 1) for k, v in kv.items()}), flush=True)
