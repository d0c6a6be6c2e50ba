"""CPU-only: the C oracle (oracle/vrf_oracle.c) against every golden vector.

`*_upstream.json` / `p256_rfc9381.json` hold values of the upstream crate's own vector files and of
RFC 9381 Appendix B: they PIN the oracle.  `*_regression.json` were produced by
the independent Python model (oracle/pyref.py, tools/gen_golden.py): they cross-check two
implementations on suites / inputs that have no upstream vector.
"""
import hashlib
import hmac
import json
import os

import numpy as np
import pytest

import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
SUITE_ID = {"bandersnatch": 0, "ed25519": 1, "secp256r1": 2}


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def hx(a):
    return np.asarray(a).tobytes().hex()


def sc_le(suite, hexstr):
    """golden scalars are in suite encoding (BE for SEC1); the ABI takes little-endian"""
    b = bytes.fromhex(hexstr)
    return b[::-1] if suite == 2 else b


def sc_enc(suite, le_arr):
    b = np.asarray(le_arr).tobytes()
    return (b[::-1] if suite == 2 else b).hex()


def test_sha2_hmac_known_answers():
    for n in (0, 1, 3, 55, 56, 63, 64, 65, 111, 112, 119, 120, 127, 128, 129, 255, 256, 1000):
        m = (bytes(range(256)) * 4)[:n]
        assert O.sha512(m) == hashlib.sha512(m).digest()
        assert O.sha256(m) == hashlib.sha256(m).digest()
    for kl in (0, 1, 32, 64, 65, 100):
        k = bytes(range(kl))
        assert O.hmac_sha256(k, b"The quick brown fox") == hmac.new(k, b"The quick brown fox", "sha256").digest()


def test_bandersnatch_upstream_ietf():
    g = load("bandersnatch_upstream.json")
    for v in g["ietf"]:
        sk, pk = O.secret_from_seed(0, [bytes.fromhex(v["seed"])])
        assert hx(sk) == v["sk"]
        assert hx(O.point_encode(0, pk)) == v["pk"]
        I, ok = O.data_to_point(0, [bytes.fromhex(v["salt"]) + bytes.fromhex(v["alpha"])])
        assert ok[0] == 1 and hx(O.point_encode(0, I)) == v["h"]
        out = O.output(0, sk, I)
        if "gamma" in v:
            assert hx(O.point_encode(0, out)) == v["gamma"]
            assert hx(O.point_to_hash(0, out)) == v["beta"]
        ad = [bytes.fromhex(v["ad"])]
        c, s = O.ietf_prove(0, sk, I, out, ad)
        if "proof_c" in v:
            assert hx(c) == v["proof_c"] and hx(s) == v["proof_s"]
        assert O.ietf_verify(0, pk, I, out, c, s, ad)[0] == 1


def test_bandersnatch_upstream_pedersen():
    g = load("bandersnatch_upstream.json")
    for v in g["pedersen"]:
        sk, pk = O.secret_from_seed(0, [bytes.fromhex(v["seed"])])
        I, _ = O.data_to_point(0, [bytes.fromhex(v["alpha"])])
        out = O.output(0, sk, I)
        ad = [bytes.fromhex(v["ad"])]
        proof, bl = O.pedersen_prove(0, sk, I, out, ad)
        assert hx(bl) == v["blinding"]
        p = proof[0]
        assert hx(O.point_encode(0, p[0:64])) == v["proof_pk_com"]
        assert hx(O.point_encode(0, p[64:128])) == v["proof_r"]
        assert hx(O.point_encode(0, p[128:192])) == v["proof_ok"]
        assert hx(p[192:224]) == v["proof_s"] and hx(p[224:256]) == v["proof_sb"]
        assert O.pedersen_verify(0, I, out, proof, ad)[0] == 1


def test_p256_rfc9381_examples():
    g = load("p256_rfc9381.json")
    for v in g["ietf"]:
        sk = np.frombuffer(sc_le(2, v["sk"]), np.uint8)
        pk_pt, ok = O.point_decode(2, bytes.fromhex(v["pk"]))
        assert ok[0] == 1
        # pk = sk*G
        G = bytes.fromhex("6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296")[::-1] + \
            bytes.fromhex("4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5")[::-1]
        assert hx(O.output(2, sk, G)) == hx(pk_pt)
        I, ok = O.data_to_point(2, [bytes.fromhex(v["pk"]) + bytes.fromhex(v["alpha"])])
        assert ok[0] == 1
        if "h" in v:
            assert hx(O.point_encode(2, I)) == v["h"]
        if "k" in v:
            assert sc_enc(2, O.nonce(2, sk, I)) == v["k"]
        out = O.output(2, sk, I)
        c, s = O.ietf_prove(2, sk, I, out, [b""])
        pi = O.point_encode(2, out).tobytes() + c.tobytes()[::-1][16:] + s.tobytes()[::-1]
        assert pi.hex() == v["pi"]
        if "beta" in v:
            assert hx(O.point_to_hash(2, out)) == v["beta"]
        assert O.ietf_verify(2, pk_pt, I, out, c, s, [b""])[0] == 1


@pytest.mark.parametrize("fname", ["bandersnatch_regression.json", "ed25519_regression.json", "p256_regression.json"])
def test_regression_vectors_batch(fname):
    g = load(fname)
    suite = SUITE_ID[g["suite"]]
    vs = g["vectors"]
    seeds = [bytes.fromhex(v["seed"]) for v in vs]
    ads = [bytes.fromhex(v["ad"]) for v in vs]
    sk, pk = O.secret_from_seed(suite, seeds)
    L = O.lib().oracle_point_enc_len(suite)
    assert [sc_enc(suite, x) for x in sk] == [v["sk"] for v in vs]
    assert [hx(x) for x in O.point_encode(suite, pk)] == [v["pk"] for v in vs]
    dec, ok = O.point_decode(suite, b"".join(bytes.fromhex(v["pk"]) for v in vs))
    assert ok.all() and (dec == pk).all()
    I, ok = O.data_to_point(suite, [bytes.fromhex(v["salt"]) + bytes.fromhex(v["alpha"]) for v in vs])
    assert ok.all()
    assert [hx(x) for x in O.point_encode(suite, I)] == [v["h"] for v in vs]
    out = O.output(suite, sk, I)
    assert [hx(x) for x in O.point_encode(suite, out)] == [v["gamma"] for v in vs]
    assert [hx(x) for x in O.point_to_hash(suite, out)] == [v["beta"] for v in vs]
    assert [sc_enc(suite, x) for x in O.nonce(suite, sk, I)] == [v["nonce"] for v in vs]
    c, s = O.ietf_prove(suite, sk, I, out, ads)
    assert [sc_enc(suite, x) for x in c] == [v["proof_c"] for v in vs]
    assert [sc_enc(suite, x) for x in s] == [v["proof_s"] for v in vs]
    assert O.ietf_verify(suite, pk, I, out, c, s, ads).all()
    proof, bl = O.pedersen_prove(suite, sk, I, out, ads)
    assert [sc_enc(suite, x) for x in bl] == [v["blinding"] for v in vs]
    for p, v in zip(proof, vs):
        enc = O.point_encode(suite, p[:192])
        assert [hx(e) for e in enc] == [v["ped_pk_com"], v["ped_r"], v["ped_ok"]]
        assert sc_enc(suite, p[192:224]) == v["ped_s"] and sc_enc(suite, p[224:256]) == v["ped_sb"]
    assert O.pedersen_verify(suite, I, out, proof, ads).all()
    # negative cases: every tampered field must be rejected
    for field in range(6):
        cc, ss, oo, pp, aa = c.copy(), s.copy(), out.copy(), pk.copy(), list(ads)
        if field == 0: cc[:, 0] ^= 1
        if field == 1: ss[:, 3] ^= 0x10
        if field == 2: oo = np.roll(out, 1, axis=0)
        if field == 3: pp = np.roll(pk, 1, axis=0)
        if field == 4: aa = [a + b"x" for a in ads]
        if field == 5: cc, ss = ss, cc
        assert not O.ietf_verify(suite, pp, I, oo, cc, ss, aa).any()
    bad = proof.copy(); bad[:, 200] ^= 1
    assert not O.pedersen_verify(suite, I, out, bad, ads).any()
    bad = proof.copy(); bad[:, :64] = proof[:, 64:128]
    assert not O.pedersen_verify(suite, I, out, bad, ads).any()


def test_verify_rejects_malformed_points():
    g = load("bandersnatch_regression.json")
    v = g["vectors"][0]
    sk, pk = O.secret_from_seed(0, [bytes.fromhex(v["seed"])])
    I, _ = O.data_to_point(0, [bytes.fromhex(v["alpha"])])
    out = O.output(0, sk, I)
    c, s = O.ietf_prove(0, sk, I, out, [bytes.fromhex(v["ad"])])
    off = out.copy(); off[0, 0] ^= 1                    # off-curve output
    assert O.ietf_verify(0, pk, I, off, c, s, [bytes.fromhex(v["ad"])])[0] == 0
    big = out.copy(); big[0, :32] = 0xFF                # non-canonical coordinate
    assert O.ietf_verify(0, pk, I, big, c, s, [bytes.fromhex(v["ad"])])[0] == 0


def test_msm_g1_regression():
    g = load("msm_g1_regression.json")
    for case in g["cases"]:
        n = case["n"]
        bases = b"".join(bytes.fromhex(b[:96])[::-1] + bytes.fromhex(b[96:])[::-1] for b in case["bases"])
        scalars = b"".join(bytes.fromhex(s)[::-1] for s in case["scalars"])
        res = O.msm_g1(bases, scalars)[0].tobytes()
        exp = bytes(96) if case["result"] is None else bytes.fromhex(case["result"][:96])[::-1] + bytes.fromhex(case["result"][96:])[::-1]
        assert res == exp, n


def test_msm_g1_matches_sum_of_generator_multiples():
    # bases = t_j * G, so MSM = (sum s_j t_j) * G : a size-independent check at a larger n
    r = int("73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001", 16)
    n = 300
    ts = [int.from_bytes(hashlib.sha512(b"t" + j.to_bytes(4, "little")).digest(), "little") % r for j in range(n)]
    ss = [int.from_bytes(hashlib.sha512(b"s" + j.to_bytes(4, "little")).digest(), "little") % r for j in range(n)]
    bases = O.g1_mul_gen(b"".join(t.to_bytes(32, "little") for t in ts))
    cols = b"".join(s.to_bytes(32, "little") for s in ss) + b"".join(((s * 7 + 1) % r).to_bytes(32, "little") for s in ss)
    got = O.msm_g1(bases, cols, n_columns=2)
    e0 = sum(s * t for s, t in zip(ss, ts)) % r
    e1 = sum(((s * 7 + 1) % r) * t for s, t in zip(ss, ts)) % r
    exp = O.g1_mul_gen(e0.to_bytes(32, "little") + e1.to_bytes(32, "little"))
    assert (got == exp).all()
