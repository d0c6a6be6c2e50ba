"""GPU parity for every per-suite entry point of the C ABI (Bandersnatch, Ed25519, secp256r1) against the CPU
oracle and the golden vectors: Secret / Public / Input / Output, codec, nonce, ietf prove, pedersen."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
import vectors as V

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TE_SUITES = [O.BANDERSNATCH, O.ED25519, O.P256, O.BANDERSNATCH_SW, O.JUBJUB, O.BABYJUBJUB]      # every suite of the crate (the last three)


@pytest.fixture(scope="module")
def eng():
    import ark_ec_vrfs_b200 as vrfs
    e = vrfs.Engine(0)
    yield e
    e.close()


def hx(a):
    return np.asarray(a).tobytes().hex()


@pytest.mark.parametrize("suite", TE_SUITES)
def test_keys_inputs_outputs(eng, suite):
    n = 300
    seeds = [bytes([i & 0xFF]) * (i % 70) for i in range(n)]          # includes the empty seed and multi-block seeds
    sk_o, pk_o = O.secret_from_seed(suite, seeds)
    sk, pk = eng.secret_from_seed(suite, seeds)
    assert np.array_equal(sk, sk_o) and np.array_equal(pk, pk_o)
    datas = [bytes((i * 3 + j) & 0xFF for j in range((i * 5) % 260)) for i in range(n)]
    inp_o, ok_o = O.data_to_point(suite, datas)
    inp, ok = eng.data_to_point(suite, datas)
    assert np.array_equal(ok, ok_o) and ok.all() and np.array_equal(inp, inp_o)
    out_o = O.output(suite, sk_o, inp_o)
    assert np.array_equal(eng.output(suite, sk_o, inp_o), out_o)
    assert np.array_equal(eng.point_to_hash(suite, out_o), O.point_to_hash(suite, out_o))
    assert np.array_equal(eng.nonce(suite, sk_o, inp_o), O.nonce(suite, sk_o, inp_o))
    enc_o = O.point_encode(suite, out_o)
    assert np.array_equal(eng.point_encode(suite, out_o), enc_o)
    dec, dok = eng.point_decode(suite, enc_o)
    assert dok.all() and np.array_equal(dec, out_o)
    # decoding arbitrary bytes: same accept/reject decisions and same points as the oracle
    L = eng.point_enc_len(suite)
    rnd = np.frombuffer(b"".join(O.sha512(b"dec%d" % i) for i in range(n)), np.uint8).reshape(n, 64)[:, :L].copy()
    if suite == O.P256:
        rnd[:, 0] = 2 + (rnd[:, 0] & 1)
        rnd[::7, 0] = 4                      # invalid SEC1 tag
    d_o, k_o = O.point_decode(suite, rnd)
    d_g, k_g = eng.point_decode(suite, rnd)
    assert np.array_equal(k_g, k_o) and 0 < k_o.sum() < n and np.array_equal(d_g, d_o)


@pytest.mark.parametrize("suite", TE_SUITES)
@pytest.mark.parametrize("ad_kind", ["empty", "ragged"])
def test_ietf_prove_matches_oracle(eng, suite, ad_kind):
    n = 300
    sk, pk, inp, out = V.make_keys_inputs(suite, n)
    ads = V.make_ads(n, ad_kind)
    c_o, s_o = O.ietf_prove(suite, sk, inp, out, ads)
    c, s = eng.ietf_prove(suite, sk, inp, out, ads)
    assert np.array_equal(c, c_o) and np.array_equal(s, s_o)
    assert eng.ietf_verify(suite, pk, inp, out, c, s, ads).all()


@pytest.mark.parametrize("suite", TE_SUITES)
@pytest.mark.parametrize("ad_kind", ["empty", "ragged"])
def test_pedersen_matches_oracle(eng, suite, ad_kind):
    n = 240
    sk, pk, inp, out = V.make_keys_inputs(suite, n)
    ads = V.make_ads(n, ad_kind)
    pr_o, bl_o = O.pedersen_prove(suite, sk, inp, out, ads)
    pr, bl = eng.pedersen_prove(suite, sk, inp, out, ads)
    assert np.array_equal(bl, bl_o) and np.array_equal(pr, pr_o)
    pr = pr.copy(); out = out.copy(); inp = inp.copy()
    for i in range(0, n, 3):
        kind = (i // 3) % 7
        if kind < 3:
            pr[i, 64 * kind : 64 * kind + 64] = pr[(i + 1) % n, 64 * kind : 64 * kind + 64]   # a valid but wrong point
        elif kind == 3:
            pr[i, 192 + (i % 30)] ^= 4          # s
        elif kind == 4:
            pr[i, 224 + (i % 30)] ^= 4          # sb
        elif kind == 5:
            out[i] = out[(i + 1) % n]
        else:
            pr[i, 70] ^= 1                      # R off the curve -> InvalidData
    exp = O.pedersen_verify(suite, inp, out, pr, ads)
    got = eng.pedersen_verify(suite, inp, out, pr, ads)
    assert 0 < exp.sum() < n and np.array_equal(got, exp)


def test_upstream_bandersnatch_vectors_through_gpu(eng):
    """every pinned field of the upstream vectors, produced by the CUDA path"""
    with open(os.path.join(GOLDEN, "bandersnatch_upstream.json")) as f:
        g = json.load(f)
    for v in g["ietf"]:
        sk, pk = eng.secret_from_seed(0, [bytes.fromhex(v["seed"])])
        assert hx(sk) == v["sk"] and hx(eng.point_encode(0, pk)) == v["pk"]
        inp, ok = eng.data_to_point(0, [bytes.fromhex(v["salt"]) + bytes.fromhex(v["alpha"])])
        assert ok.all() and hx(eng.point_encode(0, inp)) == v["h"]
        out = eng.output(0, sk, inp)
        if "gamma" in v:
            assert hx(eng.point_encode(0, out)) == v["gamma"]
            assert hx(eng.point_to_hash(0, out)) == v["beta"]
        if "proof_c" in v:
            c, s = eng.ietf_prove(0, sk, inp, out, [bytes.fromhex(v["ad"])])
            assert hx(c) == v["proof_c"] and hx(s) == v["proof_s"]
    for v in g["pedersen"]:
        sk, pk = eng.secret_from_seed(0, [bytes.fromhex(v["seed"])])
        inp, _ = eng.data_to_point(0, [bytes.fromhex(v["salt"]) + bytes.fromhex(v["alpha"])])
        out = eng.output(0, sk, inp)
        pr, bl = eng.pedersen_prove(0, sk, inp, out, [bytes.fromhex(v["ad"])])
        assert hx(bl) == v["blinding"]
        enc = eng.point_encode(0, pr[0, :192].reshape(3, 64))
        assert hx(enc[0]) == v["proof_pk_com"] and hx(enc[1]) == v["proof_r"] and hx(enc[2]) == v["proof_ok"]
        assert hx(pr[0, 192:224]) == v["proof_s"] and hx(pr[0, 224:256]) == v["proof_sb"]
        assert eng.pedersen_verify(0, inp, out, pr, [bytes.fromhex(v["ad"])]).tolist() == [1]


@pytest.mark.parametrize("fname,suite", [("bandersnatch_regression.json", 0), ("ed25519_regression.json", 1), ("p256_regression.json", 2)])
def test_regression_vectors_through_gpu(eng, fname, suite):
    """tests/golden/*_regression.json (generated by the independent Python model; Ed25519 has no upstream vector)"""
    with open(os.path.join(GOLDEN, fname)) as f:
        vs = json.load(f)["vectors"]
    seeds = [bytes.fromhex(v["seed"]) for v in vs]
    datas = [bytes.fromhex(v["salt"]) + bytes.fromhex(v["alpha"]) for v in vs]
    ads = [bytes.fromhex(v["ad"]) for v in vs]
    sk, pk = eng.secret_from_seed(suite, seeds)
    inp, ok = eng.data_to_point(suite, datas)
    assert ok.all()
    out = eng.output(suite, sk, inp)
    beta = eng.point_to_hash(suite, out)
    k = eng.nonce(suite, sk, inp)
    c, s = eng.ietf_prove(suite, sk, inp, out, ads)
    pr, bl = eng.pedersen_prove(suite, sk, inp, out, ads)
    e_pk, e_in, e_out = eng.point_encode(suite, pk), eng.point_encode(suite, inp), eng.point_encode(suite, out)
    e_pr = eng.point_encode(suite, pr[:, :192].reshape(-1, 64)).reshape(len(vs), 3, -1)
    sc = (lambda a: hx(np.asarray(a)[::-1])) if suite == 2 else hx     # golden scalars are in suite encoding (BE for SEC1)
    for i, v in enumerate(vs):
        assert sc(sk[i]) == v["sk"] and hx(e_pk[i]) == v["pk"] and hx(e_in[i]) == v["h"] and hx(e_out[i]) == v["gamma"], v["comment"]
        assert hx(beta[i]) == v["beta"] and sc(k[i]) == v["nonce"], v["comment"]
        assert sc(c[i]) == v["proof_c"] and sc(s[i]) == v["proof_s"], v["comment"]
        assert sc(bl[i]) == v["blinding"] and hx(e_pr[i, 0]) == v["ped_pk_com"] and hx(e_pr[i, 1]) == v["ped_r"] and hx(e_pr[i, 2]) == v["ped_ok"], v["comment"]
        assert sc(pr[i, 192:224]) == v["ped_s"] and sc(pr[i, 224:256]) == v["ped_sb"], v["comment"]
    assert eng.ietf_verify(suite, pk, inp, out, c, s, ads).all()
    assert eng.pedersen_verify(suite, inp, out, pr, ads).all()


def test_rfc9381_p256_examples_through_gpu(eng):
    """RFC 9381 Appendix B Examples 10-11: pi = enc(Gamma) || BE16(c) || BE32(s), produced by the CUDA path"""
    with open(os.path.join(GOLDEN, "p256_rfc9381.json")) as f:
        g = json.load(f)
    vs = g["vectors"] if "vectors" in g else g["ietf"]
    for v in vs:
        sk = np.frombuffer(bytes.fromhex(v["sk"])[::-1], np.uint8).reshape(1, 32)
        pk_enc = bytes.fromhex(v["pk"])
        pk, ok = eng.point_decode(2, pk_enc)
        assert ok.all()
        inp, ok = eng.data_to_point(2, [pk_enc + bytes.fromhex(v["alpha"])])
        assert ok.all()
        if "h" in v:
            assert hx(eng.point_encode(2, inp)) == v["h"]
        if "k" in v:
            assert hx(eng.nonce(2, sk, inp)[0][::-1]) == v["k"]
        out = eng.output(2, sk, inp)
        c, s = eng.ietf_prove(2, sk, inp, out, [b""])
        pi = hx(eng.point_encode(2, out)) + hx(c[0, :16][::-1]) + hx(s[0][::-1])
        assert pi == v["pi"], v["comment"]
        if "beta" in v:
            assert hx(eng.point_to_hash(2, out)) == v["beta"]
        assert eng.ietf_verify(2, pk, inp, out, c, s, [b""]).tolist() == [1]


def test_p256_identity_points_are_rejected(eng):
    w = V.make_ietf_proofs(O.P256, 6, "empty", corrupt=False)
    pk, inp, out, c, s = (w[k].copy() for k in ("pk", "inp", "out", "c", "s"))
    pk[0] = 0; inp[1] = 0; out[2] = 0
    c[3] = 0; s[3] = 0                                   # U = V = identity
    exp = O.ietf_verify(O.P256, pk, inp, out, c, s)
    got = eng.ietf_verify(O.P256, pk, inp, out, c, s)
    assert np.array_equal(got, exp) and exp.tolist() == [0, 0, 0, 0, 1, 1]
