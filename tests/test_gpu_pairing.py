"""GPU: BLS12-381 pairing products and the batched KZG opening check against the big-integer oracle
(oracle/pairing_ref.py): GT values, bilinearity, malformed points, honest and forged openings under a public tau."""
import random

import numpy as np
import pytest

import oracle_lib as O
from oracle import pairing_ref as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import ark_ec_vrfs_b200 as vrfs
    e = vrfs.Engine(0)
    yield e
    e.close()


def u8(b): return np.frombuffer(bytes(b), np.uint8)


def test_pairing_values_and_bilinearity(eng):
    rnd = random.Random(21)
    ps, qs, want = [], [], []
    for _ in range(4):
        a, b = rnd.randrange(1, P.R), rnd.randrange(1, P.R)
        p, q = P.g1_mul(a, P.G1_GEN), P.g2_mul(b, P.G2_GEN)
        ps.append(P.g1_to_bytes(p)); qs.append(P.g2_to_bytes(q)); want.append(P.f12_to_bytes(P.gt_cubed(P.pairing(p, q))))
    ps.append(bytes(96)); qs.append(P.g2_to_bytes(P.G2_GEN)); want.append(P.f12_to_bytes(P.F12_ONE))          # e(O, Q) = 1
    ok, gt = eng.pairing_products(u8(b"".join(ps)), u8(b"".join(qs)), 1, want_gt=True)
    assert ok.tolist() == [0, 0, 0, 0, 1]
    assert [g.tobytes() for g in gt] == want
    # products: e(aG1, bG2) e(-(ab)G1, G2) = 1 via the negate mask; a wrong scalar is rejected; three pairs
    g1s, g2s, masks, expect = [], [], [], []
    for i in range(6):
        a, b = rnd.randrange(1, P.R), rnd.randrange(1, P.R)
        ab = a * b + (1 if i % 3 == 2 else 0)
        g1s.append(P.g1_to_bytes(P.g1_mul(a, P.G1_GEN)) + P.g1_to_bytes(P.g1_mul(ab, P.G1_GEN)))
        g2s.append(P.g2_to_bytes(P.g2_mul(b, P.G2_GEN)) + P.g2_to_bytes(P.G2_GEN))
        masks.append(2); expect.append(0 if i % 3 == 2 else 1)
    ok = eng.pairing_products(u8(b"".join(g1s)), u8(b"".join(g2s)), 2, negate_masks=masks)
    assert ok.tolist() == expect
    a, b = rnd.randrange(1, P.R), rnd.randrange(1, P.R)
    q = P.g2_to_bytes(P.g2_mul(5, P.G2_GEN))
    three = P.g1_to_bytes(P.g1_mul(a, P.G1_GEN)) + P.g1_to_bytes(P.g1_mul(b, P.G1_GEN)) + P.g1_to_bytes(P.g1_mul(a + b, P.G1_GEN))
    assert eng.pairing_products(u8(three), u8(q * 3), 3, negate_masks=[4]).tolist() == [1]
    assert eng.pairing_products(u8(three), u8(q * 3), 3, negate_masks=[0]).tolist() == [0]


def test_two_pairing_products_one_warp_per_product(eng):
    """n_pairs = 2 and a small batch runs the lane programs of csrc/pairing_coop.cuh (one warp per product): GT values against the
    oracle, identity pairs, malformed points, and bit-equality with the one-thread-per-product kernel (taken by large batches)"""
    rnd = random.Random(22)
    g1s, g2s, masks, want_ok, want_gt = [], [], [], [], []
    for i in range(6):
        a, b, c, d = (rnd.randrange(1, P.R) for _ in range(4))
        p1, q1, p2, q2 = P.g1_mul(a, P.G1_GEN), P.g2_mul(b, P.G2_GEN), P.g1_mul(c, P.G1_GEN), P.g2_mul(d, P.G2_GEN)
        if i == 3: p2 = P.g1_mul(a * b * pow(d, -1, P.R) % P.R, P.G1_GEN)       # e(p1, q1) e(-p2, q2) = 1
        if i == 4: p1 = None                                                     # identity: that pair contributes 1
        m = 2 if i in (1, 3) else 0
        g1s.append(P.g1_to_bytes(p1) + P.g1_to_bytes(p2)); g2s.append(P.g2_to_bytes(q1) + P.g2_to_bytes(q2)); masks.append(m)
        e1 = P.pairing(p1, q1) if p1 is not None else P.F12_ONE
        e2 = P.pairing(P.g1_neg(p2) if m else p2, q2)
        gt = P.gt_cubed(P.f12_mul(e1, e2))
        want_gt.append(P.f12_to_bytes(gt)); want_ok.append(1 if gt == P.F12_ONE else 0)
    bad = bytearray(g1s[0]); bad[100] ^= 1
    g1s.append(bytes(bad)); g2s.append(g2s[0]); masks.append(0); want_ok.append(2); want_gt.append(P.f12_to_bytes(P.F12_ONE))
    ok, gt = eng.pairing_products(u8(b"".join(g1s)), u8(b"".join(g2s)), 2, negate_masks=masks, want_gt=True)
    assert ok.tolist() == want_ok and want_ok[3] == 1
    assert [g.tobytes() for g in gt] == want_gt
    # the same products tiled into a batch large enough for the one-thread kernel
    reps = 1300 // len(g1s) + 1
    ok2, gt2 = eng.pairing_products(u8(b"".join(g1s) * reps), u8(b"".join(g2s) * reps), 2, negate_masks=masks * reps, want_gt=True)
    assert ok2.tolist() == want_ok * reps
    assert all(gt2[j].tobytes() == want_gt[j % len(g1s)] for j in range(len(ok2)))


def test_pairing_rejects_malformed_points(eng):
    q = P.g2_to_bytes(P.g2_mul(7, P.G2_GEN)); g = P.g1_to_bytes(P.G1_GEN)
    bad1 = bytearray(g); bad1[3] ^= 1
    bad2 = bytearray(q); bad2[100] ^= 1
    g1 = u8(bytes(bad1) + g + b"\xff" * 96 + g)
    g2 = u8(q + bytes(bad2) + q + q)
    assert eng.pairing_products(g1, g2, 1).tolist() == [2, 2, 2, 0]


def _openings(k, seed, tau):
    """k honest openings of random degree-3 polynomials under the public tau: C_i = [p_i(tau)] G1, W_i = [(p_i(tau) - v_i)/(tau - z_i)] G1"""
    rnd = random.Random(seed)
    zs = [rnd.randrange(P.R) for _ in range(k)]; rs = [rnd.randrange(P.R) for _ in range(k)]
    cs, ws, vs = [], [], []
    for i in range(k):
        coeffs = [rnd.randrange(P.R) for _ in range(4)]
        ev = lambda x: sum(c * pow(x, j, P.R) for j, c in enumerate(coeffs)) % P.R
        pt, v = ev(tau), ev(zs[i])
        cs.append(pt); vs.append(v); ws.append((pt - v) * pow(tau - zs[i], -1, P.R) % P.R)
    sc = lambda xs: np.frombuffer(b"".join(x.to_bytes(32, "little") for x in xs), np.uint8).reshape(-1, 32)
    return O.g1_mul_gen(sc(cs)), sc(zs), sc(vs), O.g1_mul_gen(sc(ws)), sc(rs)


@pytest.mark.parametrize("k", [1, 5, 300, 4096])
def test_kzg_batch_verify(eng, k):
    tau = 0x1234567890abcdef1234567890abcdef1234567 % P.R
    g2, tau_g2 = u8(P.g2_to_bytes(P.G2_GEN)), u8(P.g2_to_bytes(P.g2_mul(tau, P.G2_GEN)))
    C, z, v, W, r = _openings(k, 100 + k, tau)
    for level in (0, 1, 2):
        assert eng.kzg_batch_verify(C, z, v, W, r, g2, tau_g2, check_points=level) == 1
    bad_v = v.copy(); bad_v[k // 2, 0] ^= 1
    assert eng.kzg_batch_verify(C, z, bad_v, W, r, g2, tau_g2) == 0
    bad_W = W.copy(); bad_W[k - 1] = O.g1_mul_gen(np.frombuffer((12345).to_bytes(32, "little"), np.uint8).reshape(1, 32))[0]
    assert eng.kzg_batch_verify(C, z, v, bad_W, r, g2, tau_g2) == 0
    assert eng.kzg_batch_verify(C, z, v, W, r, g2, u8(P.g2_to_bytes(P.g2_mul(tau + 1, P.G2_GEN)))) == 0       # a different verifier key
    off = C.copy(); off[0, 5] ^= 1
    assert eng.kzg_batch_verify(off, z, v, W, r, g2, tau_g2, check_points=1) == 2                               # off the curve
    if k <= 5:                                                                                                     # the oracle's own aggregated check agrees
        pts = lambda a: [P.g1_from_bytes(x.tobytes()) for x in a]
        ints = lambda a: [int.from_bytes(x.tobytes(), "little") for x in a]
        assert P.kzg_batch_verify(pts(C), ints(z), ints(v), pts(W), ints(r), P.G2_GEN, P.g2_mul(tau, P.G2_GEN))
        assert not P.kzg_batch_verify(pts(C), ints(z), ints(bad_v), pts(W), ints(r), P.G2_GEN, P.g2_mul(tau, P.G2_GEN))


def test_kzg_rejects_points_outside_the_subgroup(eng):
    """check_points = 2: a point on the curve but outside the prime-order subgroup is malformed (CanonicalDeserialize rejects it)"""
    tau = 987654321
    g2, tau_g2 = u8(P.g2_to_bytes(P.G2_GEN)), u8(P.g2_to_bytes(P.g2_mul(tau, P.G2_GEN)))
    C, z, v, W, r = _openings(3, 7, tau)
    x = 0
    while True:                                                   # a curve point with a cofactor component
        x += 1
        y2 = (x ** 3 + 4) % P.Q
        y = pow(y2, (P.Q + 1) // 4, P.Q)
        if y * y % P.Q == y2:
            break
    C2 = C.copy(); C2[1] = u8(P.g1_to_bytes((x, y)))
    assert eng.kzg_batch_verify(C2, z, v, W, r, g2, tau_g2, check_points=2) == 2
    assert eng.kzg_batch_verify(C2, z, v, W, r, g2, tau_g2, check_points=1) == 0
