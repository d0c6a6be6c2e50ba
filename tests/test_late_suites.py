"""CPU: the SURVEY 8(f)4 suites (bandersnatch_sw, jubjub, baby-jubjub).  Curve constants are checked as mathematics (on the curve,
prime-order generator, cofactor); the C oracle is held to the big-integer model on every entry point and to the committed
regression vectors (tests/golden/late_suites_regression.json, tools/gen_late_suite_vectors.py).  PARITY UNPINNED against the crate."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from oracle import pyref as R

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "late_suites_regression.json")
LATE = [3, 4, 5]


@pytest.mark.parametrize("sid", LATE)
def test_curve_constants(sid):
    S = R.SUITES[sid]; C = S.curve
    assert C.on_curve(C.G) and C.is_identity(C.mul(C.r, C.G)) and not C.is_identity(C.mul(C.h, C.G))
    assert C.on_curve(S.blinding_base) and C.is_identity(C.mul(C.r, S.blinding_base))
    # Hasse bound: the group order h * r is within 2 sqrt(p) of p + 1
    assert abs(C.h * C.r - (C.p + 1)) <= 2 * int(C.p ** 0.5) + 2
    if sid == 3:      # the short-Weierstrass form is the same curve as twisted-Edwards Bandersnatch: same field, order and cofactor
        assert C.p == R.BANDERSNATCH.p and C.r == R.BANDERSNATCH.r and C.h == R.BANDERSNATCH.h


def _pt(P): return P[0].to_bytes(32, "little") + P[1].to_bytes(32, "little")


@pytest.mark.parametrize("sid", LATE)
def test_c_oracle_matches_the_model_and_the_regression_vectors(sid):
    S = R.SUITES[sid]; C = S.curve
    g = json.load(open(GOLDEN))["suites"][S.name]
    assert g["suite_id"].encode() == S.suite_id
    for v in g["vectors"]:
        seed, alpha, ad = bytes.fromhex(v["seed"]), bytes.fromhex(v["alpha"]), bytes.fromhex(v["ad"])
        sk, pk = O.secret_from_seed(sid, [seed])
        inp, ok = O.data_to_point(sid, [alpha]); assert ok.all()
        out = O.output(sid, sk, inp)
        c, s = O.ietf_prove(sid, sk, inp, out, [ad])
        pr, bl = O.pedersen_prove(sid, sk, inp, out, [ad])
        enc = lambda a: O.point_encode(sid, a)[0].tobytes().hex()
        assert sk[0].tobytes().hex() == v["sk"] and enc(pk) == v["pk"] and enc(inp) == v["h"] and enc(out) == v["gamma"]
        assert O.point_to_hash(sid, out)[0].tobytes().hex() == v["beta"]
        assert c[0].tobytes().hex() == v["proof_c"] and s[0].tobytes().hex() == v["proof_s"] and bl[0].tobytes().hex() == v["blinding"]
        assert enc(pr[:, 0:64]) == v["proof_pk_com"] and enc(pr[:, 64:128]) == v["proof_r"] and enc(pr[:, 128:192]) == v["proof_ok"]
        assert pr[0, 192:224].tobytes().hex() == v["ped_s"] and pr[0, 224:256].tobytes().hex() == v["ped_sb"]
        assert O.ietf_verify(sid, pk, inp, out, c, s, [ad]).all() and O.pedersen_verify(sid, inp, out, pr, [ad]).all()
        # and the model itself, freshly evaluated
        ski = R.secret_from_seed(S, seed); I = R.data_to_point(S, alpha)
        assert _pt(C.mul(ski, C.G)) == pk[0].tobytes() and _pt(I) == inp[0].tobytes()
    # codec round trip + arbitrary bytes: same accept / reject decisions as the model
    rnd = np.frombuffer(b"".join(O.sha512(b"late%d-%d" % (sid, i)) for i in range(200)), np.uint8).reshape(200, 64)[:, :S.pt_len].copy()
    dec, okd = O.point_decode(sid, rnd)
    for i in range(200):
        P = R.dec_pt(S, rnd[i].tobytes())
        assert bool(okd[i]) == (P is not None)
        if P is not None:
            assert dec[i].tobytes() == _pt(P)
    assert 0 < okd.sum() < 200
