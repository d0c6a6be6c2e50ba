"""GPU parity for the wire-format entry points: checked point decoding, subgroup check, signing to
serialised signatures and verification of serialised keys + signatures, against the CPU oracle and the golden vectors."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from test_wire_formats import coset_encodings

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
SUITES = [O.BANDERSNATCH, O.ED25519, O.P256]
ORDER = {0: 0x1cfb69d4ca675f520cce760202687600ff8f87007419047174fd06b52876e7e1, 1: 2**252 + 27742317777372353535851937790883648493,
         2: 0xffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551}


@pytest.fixture(scope="module")
def eng():
    import ark_ec_vrfs_b200 as vrfs
    e = vrfs.Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("suite", SUITES)
def test_checked_decode_and_subgroup_check(eng, suite):
    enc = coset_encodings(suite, 400 if suite != O.ED25519 else 120)
    pts_o, ok_o = O.point_decode_checked(suite, enc)
    pts, ok = eng.point_decode_checked(suite, enc)
    assert np.array_equal(ok, ok_o) and np.array_equal(pts, pts_o) and 0 < ok_o.sum() < len(enc)
    plain, ok_plain = O.point_decode(suite, enc)              # on-curve points of every coset
    on = plain[ok_plain == 1]
    assert np.array_equal(eng.subgroup_check(suite, on), O.subgroup_check(suite, on))
    off = on.copy(); off[:, 3] ^= 0x40                         # off-curve / non-canonical -> 0
    assert np.array_equal(eng.subgroup_check(suite, off), O.subgroup_check(suite, off))


@pytest.mark.parametrize("suite", SUITES)
@pytest.mark.parametrize("ad_kind", ["none", "ragged"])
def test_sign_and_verify_wire(eng, suite, ad_kind):
    n = 256
    sk, pk = O.secret_from_seed(suite, [b"wire-sk-%d" % i for i in range(n)])
    pk_enc = O.point_encode(suite, pk)
    datas = [bytes((i * 11 + j) & 0xFF for j in range((i * 7) % 150)) for i in range(n)]
    ads = None if ad_kind == "none" else [bytes((i + j) & 0xFF for j in range((i * 13) % 140)) for i in range(n)]
    sig_o, sok_o = O.ietf_sign_wire(suite, sk, datas, ads)
    sig, sok = eng.ietf_sign_wire(suite, sk, datas, ads)
    assert sok_o.all() and np.array_equal(sok, sok_o) and np.array_equal(sig, sig_o), "signature bytes differ from the oracle"
    # corrupt a quarter of the items in every serialised field
    L = eng.point_enc_len(suite); cl = eng.challenge_len(suite)
    sig = sig.copy(); pk_enc = pk_enc.copy(); datas = list(datas)
    for i in range(0, n, 4):
        kind = (i // 4) % 8
        if kind == 0: sig[i, L + (i % cl)] ^= 0x04                     # c
        elif kind == 1: sig[i, L + cl + (i % 31)] ^= 0x20               # s
        elif kind == 2: sig[i, :L] = sig[(i + 1) % n, :L]               # a valid gamma of another item
        elif kind == 3: pk_enc[i] = pk_enc[(i + 1) % n]                 # wrong signer
        elif kind == 4: datas[i] = datas[i] + b"x"                      # wrong input
        elif kind == 5: sig[i, 1 + (i % 20)] ^= 0x80                    # gamma bytes: undecodable or another point
        elif kind == 6:                                                 # s + r: same residue, non-canonical encoding
            be = suite == O.P256
            sb = sig[i, L + cl:].tobytes(); sv = int.from_bytes(sb, "big" if be else "little") + ORDER[suite]
            if sv < 2**256:
                sig[i, L + cl:] = np.frombuffer(sv.to_bytes(32, "big" if be else "little"), np.uint8)
        elif kind == 7 and suite != O.P256:                             # key shifted by the 2-torsion point: on curve, not in the subgroup
            p = {0: 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, 1: 2**255 - 19}[suite]
            x = int.from_bytes(pk[i, :32].tobytes(), "little"); y = int.from_bytes(pk[i, 32:].tobytes(), "little")
            sh = np.frombuffer(((p - x) % p).to_bytes(32, "little") + ((p - y) % p).to_bytes(32, "little"), np.uint8)
            pk_enc[i] = O.point_encode(suite, sh.reshape(1, 64))[0]
    ok_o, beta_o = O.ietf_verify_wire(suite, pk_enc, datas, sig, ads)
    ok, beta = eng.ietf_verify_wire(suite, pk_enc, datas, sig, ads)
    assert np.array_equal(ok, ok_o) and np.array_equal(beta, beta_o)
    # untouched items accepted; corrupted ones rejected (secp256r1: kinds 6 and 7 leave the item valid - s + n overflows 32 bytes, cofactor 1)
    assert ok_o[1::4].all() and ok_o[2::4].all() and ok_o[3::4].all() and ok_o[::4].sum() <= (n // 16 if suite == O.P256 else 0) + n // 64 + 1
    assert np.array_equal(eng.ietf_verify_wire(suite, pk_enc, datas, sig, ads, want_hash=False), ok_o)


def test_wire_golden_vectors(eng):
    g = json.load(open(os.path.join(GOLDEN, "bandersnatch_upstream.json")))
    for v in g["ietf"]:
        if "proof_c" not in v:
            continue
        sk = np.frombuffer(bytes.fromhex(v["sk"]), np.uint8); data = bytes.fromhex(v["salt"]) + bytes.fromhex(v["alpha"]); ad = bytes.fromhex(v["ad"])
        sig, ok = eng.ietf_sign_wire(O.BANDERSNATCH, sk, [data], [ad])
        assert ok[0] and sig[0].tobytes().hex() == v["gamma"] + v["proof_c"] + v["proof_s"]
        okv, beta = eng.ietf_verify_wire(O.BANDERSNATCH, np.frombuffer(bytes.fromhex(v["pk"]), np.uint8), [data], sig, [ad])
        assert okv[0] and beta[0].tobytes().hex() == v["beta"]
    g = json.load(open(os.path.join(GOLDEN, "p256_rfc9381.json")))
    for v in g["ietf"]:
        pk = bytes.fromhex(v["pk"]); data = pk + bytes.fromhex(v["alpha"])
        sig, ok = eng.ietf_sign_wire(O.P256, np.frombuffer(bytes.fromhex(v["sk"])[::-1], np.uint8), [data], None)
        assert ok[0] and sig[0].tobytes().hex() == v["pi"]
        okv = eng.ietf_verify_wire(O.P256, np.frombuffer(pk, np.uint8), [data], np.frombuffer(bytes.fromhex(v["pi"]), np.uint8), None, want_hash=False)
        assert okv[0]


def test_wire_large_batch_roundtrip(eng):
    """2^16 signatures: sign -> verify accepts all; verdicts flip exactly where a byte was flipped (size-independent property)."""
    n = 1 << 16
    seeds = [i.to_bytes(4, "little") for i in range(256)]
    sk256, pk256 = eng.secret_from_seed(O.BANDERSNATCH, seeds)
    sk = np.tile(sk256, (n // 256, 1)); pk_enc = np.tile(eng.point_encode(O.BANDERSNATCH, pk256), (n // 256, 1))
    datas = [i.to_bytes(8, "little") for i in range(n)]
    sig, ok = eng.ietf_sign_wire(O.BANDERSNATCH, sk, datas, None)
    assert ok.all()
    flip = np.zeros(n, bool); flip[::37] = True
    sig = sig.copy(); sig[flip, 70] ^= 1
    okv, beta = eng.ietf_verify_wire(O.BANDERSNATCH, pk_enc, datas, sig, None)
    assert np.array_equal(okv == 1, ~flip)
    assert not beta[flip].any() and beta[~flip].any(axis=1).all()
    sub = np.arange(0, n, 997)                                          # oracle spot check
    ok_o, beta_o = O.ietf_verify_wire(O.BANDERSNATCH, pk_enc[sub], [datas[i] for i in sub], sig[sub], None)
    assert np.array_equal(ok_o, okv[sub]) and np.array_equal(beta_o, beta[sub])


@pytest.mark.parametrize("suite", SUITES)
def test_pedersen_sign_and_verify_wire(eng, suite):
    n = 192
    sk, pk = O.secret_from_seed(suite, [b"ped-wire-%d" % i for i in range(n)])
    datas = [bytes((i * 5 + j) & 0xFF for j in range((i * 3) % 100)) for i in range(n)]
    ads = [bytes((i + 2 * j) & 0xFF for j in range((i * 11) % 90)) for i in range(n)]
    sig_o, bl_o, ok_o = O.pedersen_sign_wire(suite, sk, datas, ads)
    sig, bl, ok = eng.pedersen_sign_wire(suite, sk, datas, ads)
    assert ok_o.all() and np.array_equal(ok, ok_o) and np.array_equal(sig, sig_o) and np.array_equal(bl, bl_o)
    L = eng.point_enc_len(suite)
    sig = sig.copy(); datas = list(datas)
    for i in range(0, n, 3):
        kind = (i // 3) % 7
        if kind < 4: sig[i, kind * L + 1 + (i % 20)] ^= 0x10          # one of the four encoded points
        elif kind == 4: sig[i, 4 * L + (i % 31)] ^= 0x08               # s
        elif kind == 5: sig[i, 4 * L + 32 + (i % 31)] ^= 0x08          # sb
        else: datas[i] = datas[i] + b"!"                               # wrong input
    exp = O.pedersen_verify_wire(suite, datas, sig, ads)
    got = eng.pedersen_verify_wire(suite, datas, sig, ads)
    assert np.array_equal(got, exp) and exp[1::3].all() and exp[2::3].all() and not exp[::3].all()
    # upstream Pedersen vector 1 through the wire form
    if suite == O.BANDERSNATCH:
        g = json.load(open(os.path.join(GOLDEN, "bandersnatch_upstream.json")))
        for v in g["pedersen"]:
            skv = np.frombuffer(bytes.fromhex(v["sk"]), np.uint8); data = bytes.fromhex(v["salt"]) + bytes.fromhex(v["alpha"]); ad = bytes.fromhex(v["ad"])
            s1, b1, k1 = eng.pedersen_sign_wire(suite, skv, [data], [ad])
            assert k1[0] and b1[0].tobytes().hex() == v["blinding"]
            assert s1[0].tobytes().hex() == v["gamma"] + v["proof_pk_com"] + v["proof_r"] + v["proof_ok"] + v["proof_s"] + v["proof_sb"]
            assert eng.pedersen_verify_wire(suite, [data], s1, [ad])[0] == 1


def test_wire_edge_cases(eng):
    import ark_ec_vrfs_b200 as vrfs
    s = O.BANDERSNATCH
    # empty batches are accepted and return empty results
    sig, ok = eng.ietf_sign_wire(s, np.zeros((0, 32), np.uint8), [])
    assert sig.shape == (0, 96) and ok.shape == (0,)
    okv, beta = eng.ietf_verify_wire(s, np.zeros((0, 32), np.uint8), [], np.zeros((0, 96), np.uint8))
    assert okv.shape == (0,) and beta.shape == (0, 64)
    assert eng.pedersen_verify_wire(s, [], np.zeros((0, 192), np.uint8)).shape == (0,)
    # empty VRF input data, empty and long additional data in one batch
    sk, pk = O.secret_from_seed(s, [b"e0", b"e1", b"e2"])
    datas = [b"", b"x", b""]; ads = [b"", b"y" * 300, b"z"]
    sig, ok = eng.ietf_sign_wire(s, sk, datas, ads)
    sig_o, _ = O.ietf_sign_wire(s, sk, datas, ads)
    assert ok.all() and np.array_equal(sig, sig_o)
    assert eng.ietf_verify_wire(s, O.point_encode(s, pk), datas, sig, ads, want_hash=False).all()
    # all-zero and all-ones keys / signatures are rejected item by item, never crash the call
    junk = np.zeros((4, 96), np.uint8); junk[1] = 0xFF; junk[2, 31] = 0x80; junk[3, :32] = sig[0, :32]
    keys = np.zeros((4, 32), np.uint8); keys[1] = 0xFF; keys[3] = O.point_encode(s, pk)[0]
    got = eng.ietf_verify_wire(s, keys, [b"a"] * 4, junk, None, want_hash=False)
    exp = O.ietf_verify_wire(s, keys, [b"a"] * 4, junk, None, want_hash=False)
    assert np.array_equal(got, exp) and not got.any()
    # decreasing offsets are a caller bug: status code, not a verdict
    bad_off = (np.zeros(8, np.uint8), np.array([0, 4, 2, 8], np.uint64))
    with pytest.raises(vrfs.VrfsError):
        eng.ietf_sign_wire(s, sk, bad_off)
    with pytest.raises(vrfs.VrfsError):
        eng.ietf_verify_wire(s, O.point_encode(s, pk), bad_off, sig)


def test_full_size_roundtrips(eng):
    """BASELINE-size batches (2^20 IETF, 2^18 Pedersen) through size-independent properties: everything the signer produces
    verifies, a flipped byte flips exactly that verdict, and a strided sample agrees with the oracle bit for bit."""
    s = O.BANDERSNATCH
    n = 1 << 20
    sk256, pk256 = eng.secret_from_seed(s, [b"full-%d" % i for i in range(256)])
    sk = np.tile(sk256, (n // 256, 1)); pk_enc = np.tile(eng.point_encode(s, pk256), (n // 256, 1))
    datas = (np.arange(n, dtype=np.uint64).view(np.uint8).copy(), np.arange(n + 1, dtype=np.uint64) * 8)
    sig, ok = eng.ietf_sign_wire(s, sk, datas)
    assert ok.all()
    flip = np.zeros(n, bool); flip[5::4099] = True
    sig = sig.copy(); sig[flip, 33] ^= 0x40
    okv, beta = eng.ietf_verify_wire(s, pk_enc, datas, sig)
    assert np.array_equal(okv == 1, ~flip) and not beta[flip].any()
    sub = np.arange(0, n, 16411)
    sub_datas = [int(i).to_bytes(8, "little") for i in sub]
    so, _ = O.ietf_sign_wire(s, sk[sub], sub_datas)
    good = ~flip[sub]
    assert np.array_equal(so[good], sig[sub][good])
    ok_o, beta_o = O.ietf_verify_wire(s, pk_enc[sub], sub_datas, sig[sub])
    assert np.array_equal(ok_o, okv[sub]) and np.array_equal(beta_o, beta[sub])
    # Pedersen, 2^18
    m = 1 << 18
    datas_m = (datas[0][:8 * m], datas[1][:m + 1])
    psig, bl, pok = eng.pedersen_sign_wire(s, sk[:m], datas_m)
    assert pok.all()
    pflip = np.zeros(m, bool); pflip[3::1031] = True
    psig = psig.copy(); psig[pflip, 170] ^= 0x02
    assert np.array_equal(eng.pedersen_verify_wire(s, datas_m, psig) == 1, ~pflip)
    subm = np.arange(0, m, 8209)
    po, bo, _ = O.pedersen_sign_wire(s, sk[subm], [int(i).to_bytes(8, "little") for i in subm])
    gm = ~pflip[subm]
    assert np.array_equal(po[gm], psig[subm][gm]) and np.array_equal(bo, bl[subm])
