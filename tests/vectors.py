"""Seeded synthetic workloads shared by the GPU parity tests, smoke() and bench.py.  Everything here is
produced by the CPU oracle (tests/oracle_lib.py) - test infrastructure, never the product path."""
import hashlib

import numpy as np

import oracle_lib as O


def make_keys_inputs(suite, n, tag=b"vrfs-b200"):
    seeds = [hashlib.sha256(tag + b"-sk-%d" % i).digest() for i in range(n)]
    sk, pk = O.secret_from_seed(suite, seeds)
    alphas = [tag + b"-alpha-" + i.to_bytes(8, "little") for i in range(n)]
    inp, ok = O.data_to_point(suite, alphas)
    assert ok.all()
    out = O.output(suite, sk, inp)
    return sk, pk, inp, out


def make_ads(n, kind):
    if kind == "empty":
        return None
    if kind == "fixed32":
        return [hashlib.sha256(i.to_bytes(8, "little")).digest() for i in range(n)]
    # ragged: lengths crossing the SHA-512 / SHA-256 block boundaries (test plan item 2)
    lens = [0, 1, 17, 55, 56, 63, 64, 65, 68, 69, 70, 111, 112, 127, 128, 129, 196, 197, 198, 300]
    return [bytes((i * 7 + j) & 0xFF for j in range(lens[i % len(lens)])) for i in range(n)]


def make_ietf_proofs(suite, n, ad_kind="empty", corrupt=True, tag=b"vrfs-b200"):
    """returns dict(pk, inp, out, c, s, ads, expect) ; 1/4 of the items are corrupted (c, s, output or pk)"""
    sk, pk, inp, out = make_keys_inputs(suite, n, tag)
    ads = make_ads(n, ad_kind)
    c, s = O.ietf_prove(suite, sk, inp, out, ads)
    pk, inp, out, c, s = (a.copy() for a in (pk, inp, out, c, s))
    if corrupt:
        for i in range(0, n, 4):
            kind = (i // 4) % 5
            if kind == 0:
                c[i, (i // 20) % 16] ^= 1 << (i % 8)
            elif kind == 1:
                s[i, (i // 20) % 31] ^= 1 << (i % 8)
            elif kind == 2:
                out[i] = out[(i + 1) % n]          # a valid point, wrong output
            elif kind == 3:
                pk[i] = pk[(i + 1) % n]            # a valid key, wrong signer
            else:
                inp[i, 5] ^= 0x10                  # almost surely off-curve -> Error::InvalidData
    expect = O.ietf_verify(suite, pk, inp, out, c, s, ads)
    return dict(sk=sk, pk=pk, inp=inp, out=out, c=c, s=s, ads=ads, expect=expect)


def tile(d, reps):
    """repeat a workload `reps` times (ads are dropped to None-or-tiled)"""
    o = {}
    for k, v in d.items():
        if isinstance(v, np.ndarray):
            o[k] = np.ascontiguousarray(np.tile(v, (reps,) + (1,) * (v.ndim - 1)))
        elif k == "ads":
            o[k] = None if v is None else v * reps
        else:
            o[k] = v
    return o
