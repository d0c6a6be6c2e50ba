"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol of
include/vrfs_b200.h, fails loudly without a GPU (no fallback), and the device headers - compiled for the
host by tests/host_emul - agree with big-integer arithmetic and with the oracle."""
import ctypes as C
import os
import random
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from oracle import pyref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ark_ec_vrfs_b200 import build, _lib
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from ark_ec_vrfs_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "vrfs_b200.h")).read()
    declared = set(re.findall(r"\b(vrfs_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.vrfs_abi_version() == 2
    assert [lib.vrfs_suite_challenge_len(i) for i in range(7)] == [32, 16, 16, 32, 32, 32, 0]
    assert [lib.vrfs_suite_hash_len(i) for i in range(7)] == [64, 64, 32, 64, 64, 64, 0]
    assert [lib.vrfs_suite_point_enc_len(i) for i in range(7)] == [32, 32, 33, 33, 32, 32, 0]


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ark_ec_vrfs_b200 as vrfs
    with pytest.raises(vrfs.VrfsError) as e:
        vrfs.Engine(0)
    assert "VRFS_CUDA_ERROR" in str(e.value)


def test_product_does_not_reference_oracle():
    """nothing under ark_ec_vrfs_b200/ may import, link or call the oracle"""
    for root, _, files in os.walk(os.path.join(ROOT, "ark_ec_vrfs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "vrf_oracle" not in txt and "pyref" not in txt.replace("oracle/pyref.py", ""), f


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "host_emul")], check=True)
    return C.CDLL(os.path.join(ROOT, "tests", "host_emul", "libhostemu.so"))


FIELDS = [(R.BLS_FR, 8), (R.BANDERSNATCH.r, 8), (R.ED25519.p, 8), (R.ED25519.r, 8), (R.P256.p, 8), (R.P256.r, 8), (R.BLS_FQ, 12)]


def _fop(emu, f, o, a, b, n):
    A = (C.c_uint32 * n)(*[(a >> (32 * i)) & 0xFFFFFFFF for i in range(n)])
    B = (C.c_uint32 * n)(*[(b >> (32 * i)) & 0xFFFFFFFF for i in range(n)])
    out = (C.c_uint32 * n)()
    emu.hostemu_field_op(f, o, A, B, out)
    return sum(int(out[i]) << (32 * i) for i in range(n))


@pytest.mark.parametrize("fi", range(len(FIELDS)))
def test_montgomery_field_arithmetic(emu, fi):
    p, n = FIELDS[fi]
    Rm = 1 if p in (R.ED25519.p, R.P256.p) else 1 << (32 * n); Ri = pow(Rm, -1, p)      # 2^255-19 and the P-256 prime are kept in plain residues (arith.cuh pm_fold / p256_fold)
    W = 1 << (32 * n)
    rnd = random.Random(fi)
    for t in range(200):
        a = [0, 1, p - 1, p - 2][t] if t < 4 else rnd.randrange(p)
        b = rnd.randrange(W) if t % 2 else rnd.choice([0, 1, p - 1, W - 1, p])
        assert _fop(emu, fi, 0, a, b, n) == a * b * Ri % p
        assert _fop(emu, fi, 9, a, 0, n) == a * a * Ri % p          # dedicated squaring
        b2 = rnd.randrange(p)
        assert _fop(emu, fi, 1, a, b2, n) == (a + b2) % p
        assert _fop(emu, fi, 2, a, b2, n) == (a - b2) % p
        assert _fop(emu, fi, 3, b, 0, n) == b * Rm % p
        assert _fop(emu, fi, 4, a, 0, n) == a * Ri % p
        assert _fop(emu, fi, 6, b, a, n) == (a * W + b) * Rm % p
    for t in range(4):
        a = rnd.randrange(1, p)
        assert _fop(emu, fi, 5, a * Rm % p, 0, n) == pow(a, -1, p) * Rm % p
        assert _fop(emu, fi, 7, a * Rm % p, 0, n) == (1 if R.legendre(a, p) == 1 else 0)
        assert _fop(emu, fi, 8, a * Rm % p, 0, n) == (1 if a > (p - 1) // 2 else 0) | ((a & 1) << 1)
    if n == 8 and p < (1 << 255):                       # binary Jacobi symbol (arith.cuh jacobi): subgroup checks use it on BLS12-381 Fr
        for t in range(300):
            a = [0, 1, p - 1, 2, 4][t] if t < 5 else (rnd.randrange(p) if t % 3 else rnd.randrange(1 << (8 * (t % 31) + 1)))
            assert _fop(emu, fi, 10, a * Rm % p, 0, n) == {1: 2, -1: 0, 0: 1}[R.legendre(a, p)], a


def test_glv_split_and_lincomb(emu):
    cv = R.BANDERSNATCH; r = cv.r; lam = R.BANDERSNATCH_GLV_LAMBDA
    rnd = random.Random(11)
    for k in [0, 1, r - 1, r // 2] + [rnd.randrange(r) for _ in range(300)]:
        K = (C.c_uint32 * 8)(*[(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)]); out = (C.c_uint32 * 10)()
        emu.hostemu_glv(K, out)
        k1 = sum(out[i] << (32 * i) for i in range(4)) * (-1 if out[8] else 1)
        k2 = sum(out[4 + i] << (32 * i) for i in range(4)) * (-1 if out[9] else 1)
        assert (k1 + k2 * lam - k) % r == 0 and abs(k1) < (1 << 127) * 7 // 15 and abs(k2) < (1 << 127) * 7 // 15
    le = lambda x: x.to_bytes(32, "little")
    pt = lambda P: le(P[0]) + le(P[1])
    pt = lambda P: bytes(64) if P is None else le(P[0]) + le(P[1])          # short-Weierstrass identity = 64 zero bytes
    for suite, cv in ((0, R.BANDERSNATCH), (1, R.ED25519), (2, R.P256)):
        r = cv.r
        P1 = cv.mul(rnd.randrange(r), cv.G); P2 = cv.mul(rnd.randrange(r), cv.G)
        for nv, nf in ((0, 1), (1, 0), (2, 0), (1, 1)):
            for neg in (0, 3, 5, 6):
                k1, k2, f = (rnd.choice([0, 1, r - 1, rnd.randrange(r)]) for _ in range(3))
                out = C.create_string_buffer(64)
                assert emu.hostemu_lincomb(suite, nv, nf, pt(P1), le(k1), pt(P2), le(k2), le(f), neg, out) == 1
                E = cv.identity()
                if nv >= 1: E = cv.add(E, cv.mul((-k1 if neg & 1 else k1) % r, P1))
                if nv >= 2: E = cv.add(E, cv.mul((-k2 if neg & 2 else k2) % r, P2))
                if nf >= 1: E = cv.add(E, cv.mul((-f if neg & 4 else f) % r, cv.G))
                assert out.raw == pt(E), (suite, nv, nf, neg)


@pytest.mark.parametrize("suite", [0, 1])
def test_host_emulated_ietf_verify_matches_oracle(emu, suite):
    import vectors as V
    n = 24
    w = V.make_ietf_proofs(suite, n, "ragged")
    ad, off = O.pack_var(w["ads"])
    got = np.zeros(n, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    emu.hostemu_ietf_verify(suite, C.c_size_t(n), p(w["pk"]), p(w["inp"]), p(w["out"]), p(w["c"]), p(w["s"]), p(ad), p(off), p(got))
    assert np.array_equal(got, w["expect"]) and 0 < got.sum() < n


def test_fq381_word_approximation_gcd_inverse(emu):
    """csrc/msm.cuh fq381_inv_fast (26 rounds of 30 binary-GCD steps on 62-bit approximations; the batched-affine rounds run it
    in lock-step on 32 lanes) against big-integer arithmetic; its own loop must finish (no fallback) for every non-zero input"""
    p = R.BLS_FQ; Rm = 1 << 384
    rnd = random.Random(22)
    special = [0, 1, p - 1, 2, (p + 1) // 2, 3, p - 2, 1 << 380, (1 << 380) - 1, (1 << 62) - 1, 1 << 62]
    for t in range(1500):
        a = special[t] if t < len(special) else (rnd.randrange(p) if t % 3 else rnd.randrange(1 << rnd.randrange(1, 381)))
        A = (C.c_uint32 * 12)(*[((a * Rm % p) >> (32 * i)) & 0xFFFFFFFF for i in range(12)]); out = (C.c_uint32 * 12)()
        finished = emu.hostemu_fq381_inv_fast(A, out)
        got = sum(int(out[i]) << (32 * i) for i in range(12))
        assert got == (pow(a, -1, p) * Rm % p if a else 0), a
        assert finished == (1 if a else 0), a


def test_g1_glv_split_of_msm_scalars(emu):
    """csrc/msm.cuh g1_glv_split: k = k1 + q z^2 with both halves below z^2 < 2^128 (z the BLS12-381 curve parameter), extremes included"""
    z2 = 0xd201000000010000 ** 2
    r = R.BLS_FR
    assert r == z2 * z2 - z2 + 1
    rnd = random.Random(24)
    ext = [0, 1, z2 - 1, z2, z2 + 1, r - 1, r - z2, 2 * z2 - 1, (z2 - 1) * z2, (z2 - 1) * z2 + z2 - 1]
    for t in range(3000):
        k = ext[t] if t < len(ext) else rnd.randrange(r) if t % 4 else rnd.randrange(z2) * z2 + rnd.choice([0, 1, z2 - 1, z2 - 2])
        k %= r
        K = (C.c_uint32 * 8)(*[(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)]); k1 = (C.c_uint32 * 4)(); q = (C.c_uint32 * 4)()
        emu.hostemu_g1_glv_split(K, k1, q)
        a = sum(int(k1[i]) << (32 * i) for i in range(4)); b = sum(int(q[i]) << (32 * i) for i in range(4))
        assert a == k % z2 and b == k // z2, hex(k)


def test_fq381_fused_product_difference(emu):
    """csrc/msm.cuh fq381_mul_sub2: a b - c d through ONE Montgomery reduction of a b + c (q - d), extremes included"""
    p = R.BLS_FQ; Rm = 1 << 384
    rnd = random.Random(23)
    lim = lambda x: (C.c_uint32 * 12)(*[((x * Rm % p) >> (32 * i)) & 0xFFFFFFFF for i in range(12)])
    ext = [0, 1, p - 1, p - 2, (p - 1) // 2]
    for t in range(600):
        a, b, c, d = ([ext[(t >> (2 * k)) % 5] for k in range(4)] if t < 200 else [rnd.randrange(p) for _ in range(4)])
        out = (C.c_uint32 * 12)()
        emu.hostemu_fq381_mul_sub2(lim(a), lim(b), lim(c), lim(d), out)
        assert sum(int(out[i]) << (32 * i) for i in range(12)) == (a * b - c * d) * Rm % p, (a, b, c, d)


def test_fq381_binary_euclid_inverse(emu):
    """csrc/msm.cuh fq381_inv (used on the MSM's final projective -> affine step) against big-integer arithmetic"""
    p = R.BLS_FQ; Rm = 1 << 384
    rnd = random.Random(21)
    for t in range(300):
        a = [0, 1, p - 1, 2, (p + 1) // 2][t] if t < 5 else rnd.randrange(p)
        A = (C.c_uint32 * 12)(*[((a * Rm % p) >> (32 * i)) & 0xFFFFFFFF for i in range(12)]); out = (C.c_uint32 * 12)()
        emu.hostemu_fq381_inv(A, out)
        got = sum(int(out[i]) << (32 * i) for i in range(12))
        assert got == (pow(a, -1, p) * Rm % p if a else 0), a


def test_ntt_domain_generators_match_ark_ff(emu):
    """csrc/ring.cuh: the FFT of vrfs_fr_fft_batch runs on Radix2EvaluationDomain's generator
    TWO_ADIC_ROOT_OF_UNITY^(2^(32 - log n)) with TWO_ADIC_ROOT_OF_UNITY = 7^((r-1)/2^32) (ark-bls12-381 Fr), and scales by 1/n"""
    r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    root = pow(7, (r - 1) >> 32, r)
    assert root == 10238227357739495823651030575849232062558860180284477541189508159991286009131
    out = (C.c_uint32 * 8)()
    val = lambda: sum(int(out[i]) << (32 * i) for i in range(8))
    for logn in (0, 1, 5, 11, 17, 26, 32):
        w = pow(root, 1 << (32 - logn), r)
        emu.hostemu_ntt_domain_gen(logn, 0, out); assert val() == w
        emu.hostemu_ntt_domain_gen(logn, 1, out); assert val() == pow(w, -1, r)
        assert pow(w, 1 << logn, r) == 1 and (logn == 0 or pow(w, 1 << (logn - 1), r) == r - 1)
        emu.hostemu_ntt_inv_n(logn, out); assert val() == pow(1 << logn, -1, r)


def test_g1_wire_format_on_the_host(emu):
    """csrc/ring.cuh g1_compress_one / g1_decompress_one (ark-bls12-381's zcash-format compressed G1): the generator's published
    encoding, round trips with both y signs, the infinity encoding, rejection of non-canonical / off-curve / flag-less inputs, and
    the endomorphism subgroup test against [r]P on points of the curve outside G1"""
    p = R.BLS_FQ; r = R.BLS_FR; Cv = R.BLS12_381_G1
    le96 = lambda P: (C.c_uint8 * 96)(*(P[0].to_bytes(48, "little") + P[1].to_bytes(48, "little")))
    out48 = (C.c_uint8 * 48)(); out96 = (C.c_uint8 * 96)()
    emu.hostemu_g1_compress(le96(Cv.G), out48)
    assert bytes(out48).hex() == "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
    rnd = random.Random(9)
    for t in range(6):
        P = Cv.mul(rnd.randrange(1, r), Cv.G)
        for Q in (P, Cv.neg(P)):
            emu.hostemu_g1_compress(le96(Q), out48)
            enc = bytes(out48)
            assert enc[0] & 0x80 and not enc[0] & 0x40 and bool(enc[0] & 0x20) == (Q[1] > p - Q[1])
            assert int.from_bytes(bytes([enc[0] & 0x1F]) + enc[1:], "big") == Q[0]
            assert emu.hostemu_g1_decompress(out48, 1, out96) == 1 and bytes(out96) == bytes(le96(Q))
    # infinity
    emu.hostemu_g1_compress((C.c_uint8 * 96)(), out48)
    assert bytes(out48) == bytes([0xC0]) + bytes(47)
    assert emu.hostemu_g1_decompress(out48, 1, out96) == 1 and not any(out96)
    bad = bytearray(bytes(out48)); bad[47] = 1
    assert emu.hostemu_g1_decompress((C.c_uint8 * 48)(*bad), 1, out96) == 0
    # no compression flag / x >= p / x that is no abscissa
    g = bytearray.fromhex("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")
    nf = bytearray(g); nf[0] &= 0x7F
    assert emu.hostemu_g1_decompress((C.c_uint8 * 48)(*nf), 1, out96) == 0
    big = bytearray((p + 1).to_bytes(48, "big")); big[0] |= 0x80
    assert emu.hostemu_g1_decompress((C.c_uint8 * 48)(*big), 0, out96) == 0
    x = 1
    while pow((x ** 3 + 4) % p, (p - 1) // 2, p) == 1: x += 1
    nx = bytearray(x.to_bytes(48, "big")); nx[0] |= 0x80
    assert emu.hostemu_g1_decompress((C.c_uint8 * 48)(*nx), 0, out96) == 0
    # points of E(Fq) outside the prime-order subgroup: accepted without the check, rejected with it
    outside = 0
    for x in range(1, 40):
        rhs = (x ** 3 + 4) % p
        y = pow(rhs, (p + 1) // 4, p)
        if y * y % p != rhs: continue
        in_g1 = Cv.is_identity(Cv.mul(r, (x, y)))
        e = bytearray(x.to_bytes(48, "big")); e[0] |= 0x80 | (0x20 if y > p - y else 0)
        buf = (C.c_uint8 * 48)(*e)
        assert emu.hostemu_g1_decompress(buf, 0, out96) == 1 and bytes(out96) == bytes(le96((x, y)))
        assert emu.hostemu_g1_decompress(buf, 1, out96) == (1 if in_g1 else 0)
        outside += not in_g1
    assert outside >= 5
