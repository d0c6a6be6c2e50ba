"""Wire formats without a GPU: the oracle's serialisation rules against the golden vectors, and the
engine's device code for checked point decoding (csrc/wire.cuh, compiled for the host) against the oracle's
`mul_bigint(r).is_zero()` subgroup test on points of every coset."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def coset_encodings(suite, n):
    """n encodings whose decoded points (when they decode) are spread over all cosets of the prime-order subgroup,
    plus encodings of subgroup points and of subgroup points shifted by the affine 2-torsion point (0,-1)."""
    L = O.lib().oracle_point_enc_len(suite)
    rnd = np.frombuffer(b"".join(O.sha512(b"wire-coset-%d" % i) for i in range(n)), np.uint8).reshape(n, 64)[:, :L].copy()
    if suite == O.P256:
        rnd[:, 0] = 2 + (rnd[:, 0] & 1)
    sk, pk = O.secret_from_seed(suite, [b"wire-%d" % i for i in range(16)])
    good = O.point_encode(suite, pk)
    extra = [good]
    if suite != O.P256:
        p = {O.BANDERSNATCH: 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, O.ED25519: 2**255 - 19}[suite]
        shifted = pk.copy()
        for i in range(len(pk)):                       # (x,y) + (0,-1) = (-x,-y): on the curve, outside the subgroup
            x = int.from_bytes(pk[i, :32].tobytes(), "little"); y = int.from_bytes(pk[i, 32:].tobytes(), "little")
            shifted[i] = np.frombuffer(((p - x) % p).to_bytes(32, "little") + ((p - y) % p).to_bytes(32, "little"), np.uint8)
        extra.append(O.point_encode(suite, shifted))
        ident = np.zeros((1, 64), np.uint8); ident[0, 32] = 1                     # (0,1)
        two = np.frombuffer((0).to_bytes(32, "little") + (p - 1).to_bytes(32, "little"), np.uint8).reshape(1, 64)   # (0,-1)
        extra.append(O.point_encode(suite, np.concatenate([ident, two])))
    return np.concatenate([rnd] + extra)


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "host_emul")], check=True)
    return C.CDLL(os.path.join(ROOT, "tests", "host_emul", "libhostemu.so"))


@pytest.mark.parametrize("suite", [O.BANDERSNATCH, O.ED25519, O.P256])
def test_checked_decode_device_code_matches_reference_subgroup_test(emu, suite):
    enc = coset_encodings(suite, 120 if suite != O.ED25519 else 60)
    n = len(enc)
    pts_o, ok_o = O.point_decode_checked(suite, enc)
    _, ok_plain = O.point_decode(suite, enc)
    pts = np.zeros((n, 64), np.uint8); ok = np.zeros(n, np.uint8)
    emu.hostemu_decode_checked(suite, C.c_size_t(n), enc.ctypes.data_as(C.c_void_p), pts.ctypes.data_as(C.c_void_p), ok.ctypes.data_as(C.c_void_p))
    assert np.array_equal(ok, ok_o) and np.array_equal(pts, pts_o)
    assert ok_o.sum() >= 16
    if suite != O.P256:
        assert (ok_plain & ~ok_o & 1).sum() >= 16, "the sample must contain on-curve points outside the subgroup"
        # subgroup points accepted, their (0,-1)-translates and (0,-1) itself rejected, the identity accepted
        assert ok_o[-34:-18].all() and not ok_o[-18:-2].any() and ok_o[-2] == 1 and ok_o[-1] == 0


def test_oracle_signature_bytes_match_upstream_and_rfc9381():
    g = json.load(open(os.path.join(GOLDEN, "bandersnatch_upstream.json")))
    seen = 0
    for v in g["ietf"]:
        if "proof_c" not in v:
            continue
        sk = np.frombuffer(bytes.fromhex(v["sk"]), np.uint8); data = bytes.fromhex(v["salt"]) + bytes.fromhex(v["alpha"]); ad = bytes.fromhex(v["ad"])
        sig, ok = O.ietf_sign_wire(O.BANDERSNATCH, sk, [data], [ad])
        assert ok[0] and sig[0].tobytes().hex() == v["gamma"] + v["proof_c"] + v["proof_s"]
        okv, beta = O.ietf_verify_wire(O.BANDERSNATCH, np.frombuffer(bytes.fromhex(v["pk"]), np.uint8), [data], sig, [ad])
        assert okv[0] and beta[0].tobytes().hex() == v["beta"]
        seen += 1
    assert seen >= 1
    g = json.load(open(os.path.join(GOLDEN, "p256_rfc9381.json")))
    for v in g["ietf"]:
        pk = bytes.fromhex(v["pk"]); data = pk + bytes.fromhex(v["alpha"])           # RFC 9381: salt = encoded public key
        sk_le = np.frombuffer(bytes.fromhex(v["sk"])[::-1], np.uint8)
        sig, ok = O.ietf_sign_wire(O.P256, sk_le, [data], None)
        assert ok[0] and sig[0].tobytes().hex() == v["pi"]                           # pi_string = gamma || c || s
        okv = O.ietf_verify_wire(O.P256, np.frombuffer(pk, np.uint8), [data], sig, None, want_hash=False)
        assert okv[0]


def test_oracle_pedersen_wire_matches_upstream():
    g = json.load(open(os.path.join(GOLDEN, "bandersnatch_upstream.json")))
    for v in g["pedersen"]:
        sk = np.frombuffer(bytes.fromhex(v["sk"]), np.uint8); data = bytes.fromhex(v["salt"]) + bytes.fromhex(v["alpha"]); ad = bytes.fromhex(v["ad"])
        sig, bl, ok = O.pedersen_sign_wire(O.BANDERSNATCH, sk, [data], [ad])
        assert ok[0] and bl[0].tobytes().hex() == v["blinding"]
        assert sig[0].tobytes().hex() == v["gamma"] + v["proof_pk_com"] + v["proof_r"] + v["proof_ok"] + v["proof_s"] + v["proof_sb"]
        assert O.pedersen_verify_wire(O.BANDERSNATCH, [data], sig, [ad])[0] == 1
        bad = sig.copy(); bad[0, 100] ^= 1
        assert O.pedersen_verify_wire(O.BANDERSNATCH, [data], bad, [ad])[0] == 0


@pytest.mark.parametrize("suite", [O.BANDERSNATCH, O.ED25519, O.P256])
def test_proof_bytes_roundtrip_device_code(emu, suite):
    sk, pk = O.secret_from_seed(suite, [b"rt"])
    inp, _ = O.data_to_point(suite, [b"alpha"]); out = O.output(suite, sk, inp)
    c, s = O.ietf_prove(suite, sk, inp, out, None)
    sl = O.ietf_signature_len(suite)
    sig = np.zeros(sl, np.uint8); cb = np.zeros(32, np.uint8); sb = np.zeros(32, np.uint8); ok = np.zeros(1, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    emu.hostemu_wire_roundtrip(suite, p(out), p(c), p(s), p(sig), p(cb), p(sb), p(ok))
    sig_o, _ = O.ietf_sign_wire(suite, sk, [b"alpha"], None)
    assert ok[0] == 1 and np.array_equal(sig, sig_o[0]) and np.array_equal(cb, c[0]) and np.array_equal(sb, s[0])
    # a non-canonical s (s + r, still 32 bytes) must be rejected
    r = {0: 0x1cfb69d4ca675f520cce760202687600ff8f87007419047174fd06b52876e7e1, 1: 2**252 + 27742317777372353535851937790883648493,
         2: 0xffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551}[suite]
    big = int.from_bytes(s[0].tobytes(), "little") + r
    if big < 2**256:
        s2 = np.frombuffer(big.to_bytes(32, "little"), np.uint8).copy()
        emu.hostemu_wire_roundtrip(suite, p(out), p(c), p(s2), p(sig), p(cb), p(sb), p(ok))
        assert ok[0] == 0


def test_bls_fr_sqrt_and_elligator2_device_code(emu):
    """csrc/h2c.cuh on the host: the Pohlig-Hellman square root of BLS12-381 Fr against big-integer arithmetic, and the
    inversion-free Elligator2 data_to_point against the oracle (incl. the upstream vector inputs)."""
    import random
    q = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    Rm = 1 << 256
    rnd = random.Random(3)
    nsq = 0
    for i in range(300):
        a = [0, 1, q - 1, 4, 5][i] if i < 5 else rnd.randrange(q)
        A = (C.c_uint32 * 8)(*[((a * Rm % q) >> (32 * k)) & 0xFFFFFFFF for k in range(8)]); out = (C.c_uint32 * 8)()
        ok = emu.hostemu_bls_sqrt(A, out)
        is_sq = a == 0 or pow(a, (q - 1) // 2, q) == 1
        assert bool(ok) == is_sq, a
        if is_sq:
            r = sum(int(out[k]) << (32 * k) for k in range(8)) * pow(Rm, -1, q) % q
            assert r * r % q == a
            nsq += 1
    assert 100 < nsq < 200
    datas = [b"", b"\x0a", b"sample"] + [bytes((i * 7 + j) & 0xFF for j in range(i % 90)) for i in range(120)]
    pts_o, ok_o = O.data_to_point(O.BANDERSNATCH, datas)
    assert ok_o.all()
    for d, p in zip(datas, pts_o):
        out = np.zeros(64, np.uint8)
        buf = np.frombuffer(d, np.uint8).copy() if d else np.zeros(1, np.uint8)
        emu.hostemu_band_h2c(buf.ctypes.data_as(C.c_void_p), len(d), out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(out, p), d
