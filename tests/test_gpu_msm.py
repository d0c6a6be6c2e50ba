"""GPU parity: vrfs_msm_g1_bls12_381 (Pippenger, CUDA) against the oracle's ark-ec-style MSM and against
size-independent properties (linearity, point-range splitting) at the BASELINE ring sizes."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
R_BLS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


@pytest.fixture(scope="module")
def eng():
    import ark_ec_vrfs_b200 as vrfs
    e = vrfs.Engine(0)
    yield e
    e.close()


def synth(n, ncol, tag=b"msm"):
    """bases tau^j * G for a public test-only tau, scalars from SHA-512"""
    tau = int.from_bytes(hashlib.sha512(b"vrfs-b200-bench-tau").digest(), "little") % R_BLS
    pw, t = [], 1
    for _ in range(n):
        pw.append(t.to_bytes(32, "little")); t = t * tau % R_BLS
    bases = O.g1_mul_gen(np.frombuffer(b"".join(pw), np.uint8).reshape(n, 32))
    sc = np.frombuffer(b"".join(hashlib.sha512(tag + b"%d" % j).digest()[:32] for j in range(n * ncol)), np.uint8).reshape(n * ncol, 32).copy()
    return bases, sc


@pytest.mark.parametrize("n,ncol", [(1, 1), (2, 3), (31, 1), (32, 2), (1000, 3), (2048, 3)])
def test_msm_matches_oracle(eng, n, ncol):
    bases, sc = synth(n, ncol)
    assert np.array_equal(eng.msm_g1(bases, sc, ncol), O.msm_g1(bases, sc, ncol))


def test_msm_edge_cases(eng):
    n = 64
    bases, sc = synth(n, 3)
    sc[:n] = 0                                                   # column 0: all-zero scalars -> identity (zeros)
    sc[n:2 * n] = np.frombuffer((R_BLS - 1).to_bytes(32, "little"), np.uint8)   # column 1: all r-1
    sc[2 * n:] = 0xFF                                            # column 2: scalars >= r, reduced on load
    bases[3] = 0                                                 # an identity base
    bases[7] = bases[8]                                          # repeated bases (P + P inside a bucket)
    got = eng.msm_g1(bases, sc, 3)
    exp = O.msm_g1(bases, sc, 3)
    assert np.array_equal(got, exp) and not got[0].any()
    assert np.array_equal(eng.msm_g1(np.zeros((0, 96), np.uint8), np.zeros((0, 32), np.uint8), 1), np.zeros((1, 96), np.uint8))


def test_msm_exceptional_additions_inside_a_bucket(eng):
    """the bucket accumulator uses incomplete XYZZ mixed additions: P + P, P + (-P) and 0 + P inside one bucket must be
    caught (repeated bases, negated bases and identity bases with equal scalars), stateless and prepared, all bucket widths"""
    P_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    for n in (8, 300, 3000):
        bases, sc = synth(n, 2, b"exc")
        neg = bases.copy()
        for i in range(n):
            y = int.from_bytes(bases[i, 48:].tobytes(), "little")
            neg[i, 48:] = np.frombuffer(((P_MOD - y) % P_MOD).to_bytes(48, "little"), np.uint8)
        b = bases.copy()
        b[1] = b[0]; b[2] = b[0]; b[3] = neg[0]                     # P, P, P, -P
        b[5] = 0                                                     # identity base
        b[n - 1] = neg[n - 2]                                        # a cancelling pair
        s = sc.copy()
        for col in range(2):
            s[col * n + 1] = s[col * n]; s[col * n + 2] = s[col * n]; s[col * n + 3] = s[col * n]     # same digits -> same buckets
            s[col * n + n - 1] = s[col * n + n - 2]
        exp = O.msm_g1(b, s, 2)
        assert np.array_equal(eng.msm_g1(b, s, 2), exp)
        h = eng.msm_g1_prepare(b)
        assert np.array_equal(h.msm(s, 2), exp)
        h.release()
        # everything cancels: the sum of k*P and k*(-P) over all bases is the identity
        bb = np.concatenate([bases, neg]); ss = np.concatenate([sc[:n], sc[:n]])
        assert not eng.msm_g1(bb, ss, 1).any()


def test_msm_regression_golden(eng):
    """tests/golden/msm_g1_regression.json: big-endian hex x||y computed by the independent Python model"""
    with open(os.path.join(GOLDEN, "msm_g1_regression.json")) as f:
        g = json.load(f)
    for case in g["cases"]:
        bases = b"".join(bytes.fromhex(b[:96])[::-1] + bytes.fromhex(b[96:])[::-1] for b in case["bases"])
        scalars = b"".join(bytes.fromhex(x)[::-1] for x in case["scalars"])
        res = eng.msm_g1(bases, scalars, 1)[0].tobytes()
        exp = bytes(96) if case["result"] is None else bytes.fromhex(case["result"][:96])[::-1] + bytes.fromhex(case["result"][96:])[::-1]
        assert res == exp, case["n"]


@pytest.mark.parametrize("logn", [11, 14, 17])
def test_msm_ring_sizes_split_and_linearity(eng, logn):
    """BASELINE config 5 domain sizes: (a) splitting the point range in two and folding the partials
    (the multi-GPU path) equals the one-shot MSM; (b) MSM(s1 + s2) = MSM(s1) + MSM(s2), checked through
    the 3-column call: column 2 holds s1 + s2 mod r"""
    n = 1 << logn
    rng = np.random.default_rng(logn)
    k = rng.integers(1, 2 ** 62, size=n, dtype=np.uint64)
    ks = np.zeros((n, 32), np.uint8); ks[:, :8] = k.view(np.uint8).reshape(n, 8)
    bases = O.g1_mul_gen(ks)                                                    # k_i * G, cheap 62-bit multiples
    s1 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); s1[:, 31] &= 0x3F
    s2 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); s2[:, 31] &= 0x3F
    s3 = np.zeros((n, 32), np.uint8)
    for i in range(n):
        v = (int.from_bytes(s1[i].tobytes(), "little") + int.from_bytes(s2[i].tobytes(), "little")) % R_BLS
        s3[i] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
    sc = np.concatenate([s1, s2, s3])
    full = eng.msm_g1(bases, sc, 3)
    h = n // 2
    sc_lo = np.concatenate([s1[:h], s2[:h], s3[:h]]); sc_hi = np.concatenate([s1[h:], s2[h:], s3[h:]])
    parts = np.stack([eng.msm_g1_partial(bases[:h], sc_lo, 3), eng.msm_g1_partial(bases[h:], sc_hi, 3)])
    assert np.array_equal(eng.g1_sum_partials(parts, 3), full)
    # linearity: column0 + column1 == column2, folded on the GPU from affine -> projective (Z = 1) partials
    def as_partial(aff):
        p = np.zeros((1, 144), np.uint8); p[0, :96] = aff; p[0, 96] = 1; return p
    s = eng.g1_sum_partials(np.stack([as_partial(full[0]), as_partial(full[1])]), 1)
    assert np.array_equal(s[0], full[2])
    if logn == 11:
        assert np.array_equal(full, O.msm_g1(bases, sc, 3))


@pytest.mark.parametrize("logn", [14, 16, 17])
def test_msm_known_answer_at_ring_sizes(eng, logn):
    """KNOWN ANSWER at the sizes the metric is quoted on: bases k_i * G with public 62-bit k_i, so every column's commitment is
    (sum k_i s_i mod r) * G - one generator multiplication by the oracle.  Prepared, stateless, and ring-shaped columns
    (random x, random y, a 0/1 selector with a long run of ones) whose skewed digits take the oversized-bucket path."""
    n = 1 << logn
    rng = np.random.default_rng(1000 + logn)
    k = rng.integers(1, 2 ** 62, size=n, dtype=np.uint64)
    ks = np.zeros((n, 32), np.uint8); ks[:, :8] = k.view(np.uint8).reshape(n, 8)
    bases = O.g1_mul_gen(ks)
    raw = rng.integers(0, 256, size=(2 * n, 40), dtype=np.uint8)
    vals = [int.from_bytes(r.tobytes(), "little") % R_BLS for r in raw]                  # uniform field elements
    sel = [1] * ((3 * n) // 4) + [0] * (n - (3 * n) // 4)
    cols = [vals[:n], vals[n:], sel]
    sc = np.frombuffer(b"".join(v.to_bytes(32, "little") for col in cols for v in col), np.uint8).reshape(3 * n, 32)
    kk = [int(x) for x in k]
    want = np.frombuffer(b"".join((sum(a * b for a, b in zip(kk, col)) % R_BLS).to_bytes(32, "little") for col in cols), np.uint8).reshape(3, 32)
    exp = O.g1_mul_gen(want)
    h = eng.msm_g1_prepare(bases)
    try:
        assert np.array_equal(h.msm(sc, 3), exp)
        assert np.array_equal(h.msm(sc[2 * n:], 1), exp[2:])                                # the selector column alone
    finally:
        h.release()
    assert np.array_equal(eng.msm_g1(bases, sc, 3), exp)


@pytest.mark.parametrize("logn", [8, 11, 13, 16])
def test_msm_prepared_equals_stateless_and_handles_skew(eng, logn):
    """prepared bases (RingContext analogue) give the same commitments; columns shaped like the ring's fixed
    columns (0/1 selector, mostly-equal values) exercise the oversized-bucket path"""
    n = 1 << logn
    rng = np.random.default_rng(100 + logn)
    k = rng.integers(1, 2 ** 62, size=n, dtype=np.uint64)
    ks = np.zeros((n, 32), np.uint8); ks[:, :8] = k.view(np.uint8).reshape(n, 8)
    bases = O.g1_mul_gen(ks)
    col0 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); col0[:, 31] &= 0x3F        # random
    col1 = np.zeros((n, 32), np.uint8); col1[: (3 * n) // 4, 0] = 1                        # selector 1..1 0..0
    col2 = np.tile(col0[0], (n, 1)); col2[::5] = col0[::5]                                 # 80 % identical scalars
    sc = np.concatenate([col0, col1, col2])
    full = eng.msm_g1(bases, sc, 3)
    h = eng.msm_g1_prepare(bases)
    try:
        assert np.array_equal(h.msm(sc, 3), full)
        assert np.array_equal(h.msm(col1, 1), full[1:2])
    finally:
        h.release()
    if logn <= 11:
        assert np.array_equal(full, O.msm_g1(bases, sc, 3))


def test_prepared_partials_fold_to_the_full_commitment(eng):
    """the multi-GPU form on one GPU: three point ranges, each with its own prepared SRS slice, partials folded"""
    n, ncol = 3000, 3
    bases, sc = synth(n, ncol, b"part")
    cols = sc.reshape(ncol, n, 32)
    cuts = [0, 1000, 1001, n]
    parts = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        h = eng.msm_g1_prepare(bases[lo:hi])
        parts.append(h.msm_partial(np.ascontiguousarray(cols[:, lo:hi]).reshape(-1, 32), ncol))
        h.release()
    got = eng.g1_sum_partials(np.stack(parts), ncol)
    assert np.array_equal(got, O.msm_g1(bases, sc, ncol))


@pytest.mark.parametrize("window_bits", [0, 8, 11])
def test_msm_exceptional_and_skewed_inputs(eng, window_bits):
    """the incomplete XYZZ mixed addition of the bucket accumulation handles its exceptional cases explicitly: identity operands,
    P + P, P + (-P), odd bucket sizes, empty buckets, oversized buckets and all-zero columns, against the oracle - with the plan's
    own window size and with two caller-chosen ones (vrfs_msm_g1_prepare_ex)"""
    P_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    for n in (5, 300, 3000):
        bases, sc = synth(n, 3, b"exc%d" % window_bits)
        neg = bases.copy()
        for i in range(n):
            y = int.from_bytes(bases[i, 48:].tobytes(), "little")
            neg[i, 48:] = np.frombuffer(((P_MOD - y) % P_MOD).to_bytes(48, "little"), np.uint8)
        b = bases.copy()
        b[1] = b[0]; b[2] = b[0]; b[3] = neg[0]                     # P, P, P, -P land in the same buckets
        b[4] = 0                                                     # identity base
        if n > 8:
            b[n - 1] = neg[n - 2]; b[6] = 0; b[7] = 0                # a cancelling pair, adjacent identities
        s = sc.copy()
        for col in range(2):
            s[col * n + 1] = s[col * n]; s[col * n + 2] = s[col * n]; s[col * n + 3] = s[col * n]
            s[col * n + n - 1] = s[col * n + n - 2]
            if n > 8: s[col * n + 7] = s[col * n + 6]
        s[2 * n:] = 0; s[2 * n: 2 * n + (2 * n) // 3, 0] = 1         # column 2: a 0/1 selector (one oversized bucket)
        exp = O.msm_g1(b, s, 3)
        h = eng.msm_g1_prepare(b, window_bits=window_bits)
        try:
            assert np.array_equal(h.msm(s, 3), exp)
            assert not h.msm(np.zeros((n, 32), np.uint8), 1).any()
        finally:
            h.release()
        assert np.array_equal(eng.msm_g1(b, s, 3), exp)


def test_msm_table_mode_digit_extremes(eng):
    """short SRS: vrfs_msm_g1_prepare keeps the 128 multiples of every 2^(8w) P_i (table mode) and a commitment is the sum of the
    entries the signed radix-256 digits select.  Scalars that sit on the digit boundaries (bytes 0x7f / 0x80 / 0xff, r - 1, values
    above r that the scalar decode reduces), against the oracle and against the bucket pipeline (a window hint selects it)"""
    R_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    n = 96
    bases, sc = synth(n, 3, b"table")
    special = [0, 1, R_MOD - 1, R_MOD - 2, 0x80, 0x7f, 0x81, 0xff, 0x100, 0x8000, 0x7fff, 0x8080, 1 << 254, (1 << 254) - 1,
               int.from_bytes(b"\x80" * 31 + b"\x00", "little"), int.from_bytes(b"\x7f" * 31 + b"\x00", "little"),
               int.from_bytes(b"\x80" * 31 + b"\x73", "little") % R_MOD, int.from_bytes(b"\xff" * 31 + b"\x72", "little")]
    s = sc.copy()
    for j, v in enumerate(special):
        s[j] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
        s[n + j] = np.frombuffer(((R_MOD - v) % R_MOD).to_bytes(32, "little"), np.uint8)
    s[2 * n] = 0xff                                              # 2^256 - 1: not canonical, reduced mod r like the reference's scalar decode
    s[2 * n + 1] = np.frombuffer((R_MOD + 5).to_bytes(32, "little"), np.uint8)
    exp = O.msm_g1(bases, s, 3)
    table = eng.msm_g1_prepare(bases)
    buckets = eng.msm_g1_prepare(bases, window_bits=10)
    try:
        assert np.array_equal(table.msm(s, 3), exp)
        assert np.array_equal(buckets.msm(s, 3), exp)
        assert np.array_equal(table.msm(s[:n], 1), exp[:1])
    finally:
        table.release(); buckets.release()


def test_fq381_inverse_on_the_gpu(eng):
    """fq381_inv_fast as compiled for the GPU: correct, and its word-approximation GCD finishes by itself (no fallback)"""
    P_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    import random
    rnd = random.Random(5)
    vals = [0, 1, 2, P_MOD - 1, (P_MOD + 1) // 2, 1 << 380, (1 << 62) - 1] + [rnd.randrange(P_MOD) for _ in range(2000)] + [rnd.randrange(1 << rnd.randrange(1, 381)) for _ in range(500)]
    a = np.frombuffer(b"".join(v.to_bytes(48, "little") for v in vals), np.uint8).reshape(-1, 48)
    inv, ok = eng.fq381_inv(a)
    got = [int.from_bytes(r.tobytes(), "little") for r in inv]
    assert got == [pow(v, -1, P_MOD) if v else 0 for v in vals]
    assert list(ok) == [1 if v else 0 for v in vals]
