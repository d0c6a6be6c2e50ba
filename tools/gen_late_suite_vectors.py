#!/usr/bin/env python3
"""Regression vectors for the SURVEY 8(f)4 suites (bandersnatch_sw, jubjub, baby-jubjub) from the big-integer model
oracle/pyref.py -> tests/golden/late_suites_regression.json.  These are NOT reference vectors: suite strings, CHALLENGE_LEN and the
codec are recalled, the Pedersen blinding bases are placeholders (PARITY UNPINNED); they freeze the model so that the C oracle
and the CUDA engine can be held to it."""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import pyref as R

out = {"note": "regression values computed by oracle/pyref.py; parity unpinned (see tools/gen_late_suite_vectors.py)", "suites": {}}
for sid in (3, 4, 5):
    S = R.SUITES[sid]; C = S.curve
    vecs = []
    for i, (seed, alpha, ad) in enumerate([(b"\x01", b"", b""), (b"\x02", b"\x0a", b""), (b"seed-3", b"alpha", b"\x0b\x8c"), (b"", b"x" * 200, b"y" * 70)]):
        sk = R.secret_from_seed(S, seed); pk = C.mul(sk, C.G)
        I = R.data_to_point(S, alpha); O = C.mul(sk, I)
        c, s = R.ietf_prove(S, sk, I, O, ad)
        proof, b = R.pedersen_prove(S, sk, I, O, ad)
        assert R.ietf_verify(S, pk, I, O, ad, c, s) and R.pedersen_verify(S, I, O, ad, proof)
        yb, r_, ok, ps, psb = proof
        vecs.append(dict(seed=seed.hex(), alpha=alpha.hex(), ad=ad.hex(), sk=R.enc_sc(S, sk).hex(), pk=R.enc_pt(S, pk).hex(), h=R.enc_pt(S, I).hex(),
                         gamma=R.enc_pt(S, O).hex(), beta=R.point_to_hash(S, O).hex(), proof_c=R.enc_sc(S, c).hex(), proof_s=R.enc_sc(S, s).hex(),
                         blinding=R.enc_sc(S, b).hex(), proof_pk_com=R.enc_pt(S, yb).hex(), proof_r=R.enc_pt(S, r_).hex(), proof_ok=R.enc_pt(S, ok).hex(),
                         ped_s=R.enc_sc(S, ps).hex(), ped_sb=R.enc_sc(S, psb).hex()))
    out["suites"][S.name] = dict(suite_id=S.suite_id.decode(), index=sid, blinding_base=R.enc_pt(S, S.blinding_base).hex(), vectors=vecs)
path = os.path.join(ROOT, "tests", "golden", "late_suites_regression.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)
