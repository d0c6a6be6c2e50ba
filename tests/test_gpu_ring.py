"""GPU tests of SURVEY.md 8f-2: the fixed columns of a ring, the radix-2 FFT over BLS12-381 Fr and the three KZG commitments
(vrfs_ring_fixed_columns, vrfs_fr_fft_batch, vrfs_ring_commit; api.RingContext).  The ring-proof crate is not available offline,
so these tests pin MATHEMATICS (the FFT against its definition over ark-ff's domain generator; Lagrange-basis and monomial SRS
give the same commitment, equal to [sum_i col_i L_i(tau)] G computed with big integers) - the row layout itself is a parameter."""
import hashlib

import numpy as np
import pytest

import oracle_lib as O
import vectors as V

pytestmark = pytest.mark.gpu
R_BLS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
ROOT = 10238227357739495823651030575849232062558860180284477541189508159991286009131   # ark-ff Fr::TWO_ADIC_ROOT_OF_UNITY = 7^((r-1)/2^32)


@pytest.fixture(scope="module")
def eng():
    import ark_ec_vrfs_b200 as vrfs
    e = vrfs.Engine(0)
    yield e
    e.close()


def le32(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), np.uint8).reshape(-1, 32).copy()


def ints(a):
    return [int.from_bytes(r.tobytes(), "little") for r in np.asarray(a, np.uint8).reshape(-1, 32)]


def domain_gen(logn):
    assert pow(7, (R_BLS - 1) >> 32, R_BLS) == ROOT
    return pow(ROOT, 1 << (32 - logn), R_BLS)


def rand_fr(n, tag):
    return [int.from_bytes(hashlib.sha512(tag + b"%d" % i).digest(), "little") % R_BLS for i in range(n)]


@pytest.mark.parametrize("logn", [0, 1, 4, 7])
def test_fft_matches_its_definition(eng, logn):
    n = 1 << logn
    w = domain_gen(logn)
    cols = [rand_fr(n, b"fft%d-%d" % (logn, c)) for c in range(2)]
    got = eng.fr_fft(le32(cols[0] + cols[1]), 2, inverse=False)
    for c in range(2):
        exp = [sum(cols[c][j] * pow(w, i * j, R_BLS) for j in range(n)) % R_BLS for i in range(n)]
        assert ints(got[c * n:(c + 1) * n]) == exp
    # ifft is the inverse map, natural order on both sides
    assert ints(eng.fr_fft(got, 2, inverse=True)) == cols[0] + cols[1]


@pytest.mark.parametrize("logn,ncol", [(10, 1), (11, 3), (13, 2), (17, 3)])
def test_fft_round_trip_and_linearity(eng, logn, ncol):
    n = 1 << logn
    rng = np.random.default_rng(logn)
    a = rng.integers(0, 256, size=(ncol * n, 32), dtype=np.uint8); a[:, 31] &= 0x3F
    ev = eng.fr_fft(a, ncol)
    assert np.array_equal(eng.fr_fft(ev, ncol, inverse=True), a)
    # evaluation 0 is the sum of the coefficients; a constant polynomial evaluates to itself everywhere
    for c in range(ncol):
        assert ints(ev[c * n:c * n + 1])[0] == sum(ints(a[c * n:(c + 1) * n])) % R_BLS
    const = np.zeros((n, 32), np.uint8); const[0] = a[0]
    assert np.array_equal(eng.fr_fft(const, 1), np.tile(a[0], (n, 1)))
    # values >= r are reduced on load
    big = np.full((n, 32), 0xFF, np.uint8)
    assert ints(eng.fr_fft(eng.fr_fft(big, 1), 1, inverse=True)) == [((1 << 256) - 1) % R_BLS] * n


def test_ring_fixed_columns_layout(eng):
    n, part = 64, 40
    _, pk, inp, _ = V.make_keys_inputs(O.BANDERSNATCH, 20)
    keys, tail, padding = pk[:7], inp[:10], pk[19]
    cols = eng.ring_fixed_columns(n, part, keys, padding, tail)
    assert cols.shape == (3, n, 32)
    pts = np.concatenate([cols[0], cols[1]], axis=1)                  # row i = x || y
    assert np.array_equal(pts[:7], keys)
    assert np.array_equal(pts[7:part], np.tile(padding, (part - 7, 1)))
    assert np.array_equal(pts[part:part + 10], tail)
    assert not pts[part + 10:].any()
    sel = ints(cols[2])
    assert sel == [1] * part + [0] * (n - part)
    # a full ring, no tail, no padding needed
    cols = eng.ring_fixed_columns(16, 16, pk[:16], padding, np.zeros((0, 64), np.uint8))
    assert np.array_equal(np.concatenate([cols[0], cols[1]], axis=1), pk[:16]) and ints(cols[2]) == [1] * 16
    import ark_ec_vrfs_b200 as vrfs
    with pytest.raises(vrfs.VrfsError):
        eng.ring_fixed_columns(48, 40, keys, padding, tail)            # not a power of two
    with pytest.raises(vrfs.VrfsError):
        eng.ring_fixed_columns(64, 60, keys, padding, tail)            # tail does not fit


def test_ring_commit_lagrange_equals_monomial_equals_direct(eng):
    """test-only SRS with a public tau: [tau^i]G (monomial) and [L_i(tau)]G (Lagrange over the radix-2 domain)"""
    import ark_ec_vrfs_b200 as vrfs
    logn = 8; n = 1 << logn
    tau = int.from_bytes(hashlib.sha512(b"vrfs-b200-bench-tau").digest(), "little") % R_BLS
    w = domain_gen(logn)
    mono = [pow(tau, i, R_BLS) for i in range(n)]
    zn = (pow(tau, n, R_BLS) - 1) * pow(n, -1, R_BLS) % R_BLS
    lag = [zn * pow(w, i, R_BLS) * pow((tau - pow(w, i, R_BLS)) % R_BLS, -1, R_BLS) % R_BLS for i in range(n)]
    assert sum(lag) % R_BLS == 1
    srs_mono, srs_lag = O.g1_mul_gen(le32(mono)), O.g1_mul_gen(le32(lag))
    _, pk, inp, _ = V.make_keys_inputs(O.BANDERSNATCH, 150)
    keys, tail, padding = pk[:100], inp[:50], pk[149]
    part = n - 3 - len(tail) - 1
    from ark_ec_vrfs_b200 import api
    suite = api.Suite(vrfs.BANDERSNATCH, eng)
    cols = eng.ring_fixed_columns(n, part, keys, padding, tail)
    direct = O.g1_mul_gen(le32([sum(c * l for c, l in zip(ints(cols[k]), lag)) % R_BLS for k in range(3)]))
    outs = []
    for srs, is_lag in ((srs_lag, True), (srs_mono, False)):
        h = eng.msm_g1_prepare(srs)
        try:
            outs.append(h.ring_commit(keys, part, padding, tail, lagrange=is_lag))
            if is_lag:
                assert np.array_equal(h.msm(cols.reshape(-1, 32), 3), outs[-1])
        finally:
            h.release()
    assert np.array_equal(outs[0], direct) and np.array_equal(outs[1], direct)
    # the API mirror: defaults of the row layout, serialised RingCommitment
    ctx = api.RingContext(suite, srs_lag, True, padding, tail)
    try:
        assert ctx.keyset_part_size == part and np.array_equal(ctx.verifier_key_commitment(keys), direct)
        # the incremental form (all-padding commitment + delta over the real keys) gives the same points for every ring size
        for nk in (0, 1, 100, part):
            ks = np.tile(keys, (3, 1))[:nk]
            assert np.array_equal(ctx.verifier_key_commitment(ks, incremental=True), ctx.verifier_key_commitment(ks, incremental=False)), nk
        assert np.array_equal(ctx.verifier_key_commitment(np.tile(padding, (5, 1)), incremental=True), ctx.verifier_key_commitment(keys[:0], incremental=False))
        blob = ctx.ring_commitment_bytes(keys)
        assert blob.shape == (144,) and all(blob[48 * k] & 0x80 for k in range(3)) and not any(blob[48 * k] & 0x40 for k in range(3))
        x0 = int.from_bytes(bytes([blob[0] & 0x1F]) + blob[1:48].tobytes(), "big")
        assert x0 == int.from_bytes(direct[0, :48].tobytes(), "little")
    finally:
        ctx.release()
    # the same context from the SRS as stored (compressed points, validated on the GPU); a corrupted point is refused
    enc = eng.g1_compress(srs_lag)
    ctx2 = api.RingContext.from_compressed_srs(suite, enc, True, padding, tail)
    try:
        assert np.array_equal(ctx2.verifier_key_commitment(keys), direct)
    finally:
        ctx2.release()
    enc[5, 47] ^= 1
    with pytest.raises(ValueError):
        api.RingContext.from_compressed_srs(suite, enc, True, padding, tail)


def test_ring_commit_at_ring_size_2p10(eng):
    """domain 2^11 (ring size 2^10): the one-call commitment equals the MSM of the columns it builds"""
    n = 1 << 11
    rng = np.random.default_rng(3)
    ks = np.zeros((n, 32), np.uint8); ks[:, :8] = rng.integers(1, 2 ** 62, size=n, dtype=np.uint64).view(np.uint8).reshape(n, 8)
    srs = O.g1_mul_gen(ks)
    _, pk, inp, _ = V.make_keys_inputs(O.BANDERSNATCH, 1024)
    tail = np.tile(inp[:23], (11, 1))                                   # 253 rows
    part = n - 3 - len(tail) - 1
    h = eng.msm_g1_prepare(srs)
    try:
        cols = eng.ring_fixed_columns(n, part, pk, pk[0], tail)
        got = h.ring_commit(pk, part, pk[0], tail, lagrange=True)
        assert np.array_equal(got, h.msm(cols.reshape(-1, 32), 3)) and got.any()
        coef = eng.fr_fft(cols.reshape(-1, 32), 3, inverse=True)
        assert np.array_equal(h.ring_commit(pk, part, pk[0], tail, lagrange=False), h.msm(coef, 3))
    finally:
        h.release()


def test_g1_compress_decompress_on_the_gpu(eng):
    """vrfs_g1_compress_batch / vrfs_g1_decompress_batch against the host-side twin (api.g1_compress), the published encoding of the
    generator, round trips incl. the identity, and rejection of points outside the prime-order subgroup"""
    from ark_ec_vrfs_b200 import api
    P_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    rng = np.random.default_rng(12)
    n = 1500
    ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); ks[:, 31] &= 0x3F; ks[0] = 0; ks[0, 0] = 1; ks[1] = 0       # G, identity
    pts = O.g1_mul_gen(ks)
    neg = pts.copy()
    for i in range(2, 40):                                           # both y signs
        y = int.from_bytes(pts[i, 48:].tobytes(), "little")
        neg[i, 48:] = np.frombuffer(((P_MOD - y) % P_MOD).to_bytes(48, "little"), np.uint8)
    allp = np.concatenate([pts, neg[2:40]])
    enc = eng.g1_compress(allp)
    assert np.array_equal(enc, api.g1_compress(allp))
    assert enc[0].tobytes().hex() == "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
    assert enc[1].tobytes() == bytes([0xC0]) + bytes(47)
    back, ok = eng.g1_decompress(enc)
    assert ok.all() and np.array_equal(back, allp)
    # malformed encodings and points off the subgroup
    bad = enc[:8].copy()
    bad[2, 0] &= 0x7F                                                # compression flag missing
    bad[3] = np.frombuffer((P_MOD + 5).to_bytes(48, "big"), np.uint8); bad[3, 0] |= 0x80      # x >= p
    bad[1, 47] = 7                                                   # stray bits on the infinity encoding
    x = 1; offs = []
    while len(offs) < 4:                                             # curve points with small x are (almost always) outside G1
        rhs = (x ** 3 + 4) % P_MOD; y = pow(rhs, (P_MOD + 1) // 4, P_MOD)
        if y * y % P_MOD == rhs:
            e = bytearray(x.to_bytes(48, "big")); e[0] |= 0x80 | (0x20 if y > P_MOD - y else 0); offs.append((bytes(e), x, y))
        x += 1
    for j, (e, _, _) in enumerate(offs):
        bad[4 + j] = np.frombuffer(e, np.uint8)
    out, ok = eng.g1_decompress(bad, check_subgroup=True)
    assert list(ok) == [1, 0, 0, 0, 0, 0, 0, 0] and not out[1:].any()
    out, ok = eng.g1_decompress(bad, check_subgroup=False)
    assert list(ok) == [1, 0, 0, 0, 1, 1, 1, 1]
    for j, (_, xx, yy) in enumerate(offs):
        assert out[4 + j].tobytes() == xx.to_bytes(48, "little") + yy.to_bytes(48, "little")


def test_ring_api_edge_cases_and_argument_checks(eng):
    import ark_ec_vrfs_b200 as vrfs
    _, pk, inp, _ = V.make_keys_inputs(O.BANDERSNATCH, 40)
    rng = np.random.default_rng(8)
    ks = np.zeros((64, 32), np.uint8); ks[:, :8] = rng.integers(1, 2 ** 62, size=64, dtype=np.uint64).view(np.uint8).reshape(64, 8)
    srs = O.g1_mul_gen(ks)
    none = np.zeros((0, 64), np.uint8)
    h = eng.msm_g1_prepare(srs)
    try:
        # empty ring, no tail: only the selector column is non-trivial; keyset_part_size = 0: everything is the identity
        got = h.ring_commit(none, 16, pk[0], none, lagrange=True)
        cols = eng.ring_fixed_columns(64, 16, none, pk[0], none)
        assert np.array_equal(got, O.msm_g1(srs, cols.reshape(-1, 32), 3)) and got.any()
        assert not h.ring_commit(none, 0, pk[0], none, lagrange=True).any()
        assert not h.ring_commit_delta(none, pk[0]).any()
        # a full domain of keys (no padding, no tail), both SRS kinds against the oracle MSM of the columns / coefficients
        keys = np.tile(pk, (2, 1))[:64]
        cols = eng.ring_fixed_columns(64, 64, keys, pk[0], none).reshape(-1, 32)
        assert np.array_equal(h.ring_commit(keys, 64, pk[0], none, lagrange=True), O.msm_g1(srs, cols, 3))
        assert np.array_equal(h.ring_commit(keys, 64, pk[0], none, lagrange=False), O.msm_g1(srs, eng.fr_fft(cols, 3, inverse=True), 3))
        with pytest.raises(vrfs.VrfsError):
            h.ring_commit(keys, 63, pk[0], none)                 # more keys than key slots
        with pytest.raises(vrfs.VrfsError):
            h.ring_commit(keys[:3], 60, pk[0], inp[:5])          # tail does not fit
    finally:
        h.release()
    h3 = eng.msm_g1_prepare(srs[:48])                            # an SRS that does not cover a power-of-two domain
    try:
        with pytest.raises(vrfs.VrfsError):
            h3.ring_commit(none, 16, pk[0], none)
    finally:
        h3.release()
    # empty batches are no-ops
    assert eng.g1_compress(np.zeros((0, 96), np.uint8)).shape == (0, 48)
    pts, ok = eng.g1_decompress(np.zeros((0, 48), np.uint8))
    assert pts.shape == (0, 96) and ok.shape == (0,)
    with pytest.raises((vrfs.VrfsError, AssertionError)):
        eng.fr_fft(np.zeros((3, 32), np.uint8), 1)               # not a power of two


def test_msm_32_columns_over_one_base(eng):
    """the shape bench.py uses to derive an SRS: 32 one-scalar columns over a single prepared base"""
    ks = np.zeros((1, 32), np.uint8); ks[0, 0] = 1
    gen = O.g1_mul_gen(ks)
    rng = np.random.default_rng(4)
    sc = rng.integers(0, 256, size=(32, 32), dtype=np.uint8); sc[:, 31] &= 0x3F; sc[5] = 0
    h = eng.msm_g1_prepare(gen)
    try:
        assert np.array_equal(h.msm(sc, 32), O.g1_mul_gen(sc))
    finally:
        h.release()


def test_sharded_ring_context_on_one_rank(eng):
    """dist.ShardedRingContext with world size 1 (gloo): host-built column rows + prepared partial MSM + fold = vrfs_ring_commit;
    dist.ring_column_rows equals the device column builder"""
    import torch.distributed as dist
    from ark_ec_vrfs_b200 import dist as D
    os_env = __import__("os").environ
    os_env.setdefault("MASTER_ADDR", "127.0.0.1"); os_env.setdefault("MASTER_PORT", "29541")
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        n = 256
        rng = np.random.default_rng(6)
        ks = np.zeros((n, 32), np.uint8); ks[:, :8] = rng.integers(1, 2 ** 62, size=n, dtype=np.uint64).view(np.uint8).reshape(n, 8)
        srs = O.g1_mul_gen(ks)
        _, pk, inp, _ = V.make_keys_inputs(O.BANDERSNATCH, 120)
        keys, tail, padding = pk[:100], inp[:50], pk[119]
        part = n - 3 - len(tail) - 1
        assert np.array_equal(D.ring_column_rows(0, n, part, keys, padding, tail), eng.ring_fixed_columns(n, part, keys, padding, tail))
        ctx = D.ShardedRingContext(eng, srs, part, padding, tail)
        h = eng.msm_g1_prepare(srs)
        try:
            assert np.array_equal(ctx.verifier_key_commitment(keys), h.ring_commit(keys, part, padding, tail, lagrange=True))
        finally:
            ctx.release(); h.release()
    finally:
        if created:
            dist.destroy_process_group()
