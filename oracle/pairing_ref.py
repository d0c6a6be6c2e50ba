"""TEST INFRASTRUCTURE ONLY - big-integer Python model of the BLS12-381 pairing and of the batched KZG opening check behind
`ring` -> ring-proof's verifier (names re-exported at /root/reference/src/lib.rs:13-17: `ring`, `ring_suite_types`; SURVEY.md
8(f)3).  The implementation it restates is third-party and not mounted: w3f `ring-proof` (+ its KZG crate) over `ark-bls12-381`
/ `ark-ec` 0.5 (`Bls12::multi_miller_loop`, `final_exponentiation`).  PARITY UNPINNED against the crate (no vector offline); the
model is pinned to mathematics instead: bilinearity, non-degeneracy, the curve orders, and - for the KZG check - polynomials
committed under a PUBLIC tau.

Deliberately naive and independent of the CUDA code: affine Miller loop with explicit slopes over F_q^12 built as the tower
F_q2 = F_q[u]/(u^2+1), F_q6 = F_q2[v]/(v^3 - (u+1)), F_q12 = F_q6[w]/(w^2 - v), and the final exponentiation as ONE big-integer
power (q^12 - 1)/r.  The engine uses projective line functions, sparse products, Frobenius maps and the x-chain for the hard
part; `gt_cubed` maps this model's value to the engine's (the engine's hard part carries the usual factor 3).

Encodings match include/vrfs_b200.h: F_q elements 48-byte little-endian; G1 affine x || y (96 B, zeros = identity);
G2 affine x.c0 || x.c1 || y.c0 || y.c1 (192 B, zeros = identity); GT = 12 F_q coefficients c0.c0.c0, c0.c0.c1, c0.c1.c0, ...
(F_q12 -> F_q6 (c0, c1) -> F_q2 (c0, c1, c2) -> F_q (c0, c1)), 576 B."""
from __future__ import annotations

Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
X_ABS = 0xd201000000010000        # the curve parameter is x = -X_ABS
G1_GEN = (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
          0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)
G2_GEN = ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
           0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
          (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
           0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be))


# ---- F_q2 = F_q[u]/(u^2 + 1): pairs (c0, c1) -------------------------------------------------------------------------------
def f2(a, b=0): return (a % Q, b % Q)
F2_ZERO, F2_ONE = (0, 0), (1, 0)
XI = (1, 1)                                     # the non-residue u + 1
def f2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def f2_sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
def f2_neg(a): return ((-a[0]) % Q, (-a[1]) % Q)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
def f2_sqr(a): return f2_mul(a, a)
def f2_scale(a, k): return (a[0] * k % Q, a[1] * k % Q)
def f2_conj(a): return (a[0], (-a[1]) % Q)
def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return (a[0] * d % Q, (-a[1]) * d % Q)
def f2_mul_xi(a): return ((a[0] - a[1]) % Q, (a[0] + a[1]) % Q)
def f2_pow(a, e):
    r = F2_ONE
    while e:
        if e & 1: r = f2_mul(r, a)
        a = f2_sqr(a); e >>= 1
    return r


# ---- F_q6 = F_q2[v]/(v^3 - xi): triples -------------------------------------------------------------------------------------
F6_ZERO, F6_ONE = (F2_ZERO, F2_ZERO, F2_ZERO), (F2_ONE, F2_ZERO, F2_ZERO)
def f6_add(a, b): return tuple(f2_add(x, y) for x, y in zip(a, b))
def f6_sub(a, b): return tuple(f2_sub(x, y) for x, y in zip(a, b))
def f6_neg(a): return tuple(f2_neg(x) for x in a)
def f6_mul(a, b):
    a0, a1, a2 = a; b0, b1, b2 = b
    c0 = f2_add(f2_mul(a0, b0), f2_mul_xi(f2_add(f2_mul(a1, b2), f2_mul(a2, b1))))
    c1 = f2_add(f2_add(f2_mul(a0, b1), f2_mul(a1, b0)), f2_mul_xi(f2_mul(a2, b2)))
    c2 = f2_add(f2_add(f2_mul(a0, b2), f2_mul(a1, b1)), f2_mul(a2, b0))
    return (c0, c1, c2)
def f6_mul_v(a): return (f2_mul_xi(a[2]), a[0], a[1])
def f6_inv(a):
    a0, a1, a2 = a
    t0 = f2_sub(f2_sqr(a0), f2_mul_xi(f2_mul(a1, a2)))
    t1 = f2_sub(f2_mul_xi(f2_sqr(a2)), f2_mul(a0, a1))
    t2 = f2_sub(f2_sqr(a1), f2_mul(a0, a2))
    d = f2_inv(f2_add(f2_mul(a0, t0), f2_mul_xi(f2_add(f2_mul(a2, t1), f2_mul(a1, t2)))))
    return (f2_mul(t0, d), f2_mul(t1, d), f2_mul(t2, d))


# ---- F_q12 = F_q6[w]/(w^2 - v): pairs ---------------------------------------------------------------------------------------
F12_ONE = (F6_ONE, F6_ZERO)
def f12_mul(a, b):
    a0, a1 = a; b0, b1 = b
    return (f6_add(f6_mul(a0, b0), f6_mul_v(f6_mul(a1, b1))), f6_add(f6_mul(a0, b1), f6_mul(a1, b0)))
def f12_sqr(a): return f12_mul(a, a)
def f12_conj(a): return (a[0], f6_neg(a[1]))
def f12_inv(a):
    a0, a1 = a
    d = f6_inv(f6_sub(f6_mul(a0, a0), f6_mul_v(f6_mul(a1, a1))))
    return (f6_mul(a0, d), f6_neg(f6_mul(a1, d)))
def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1: r = f12_mul(r, a)
        a = f12_sqr(a); e >>= 1
    return r
def f12_from_fq(k): return (((k % Q, 0), F2_ZERO, F2_ZERO), F6_ZERO)
def f12_coeffs(a):
    """the 12 F_q coefficients in the engine's order"""
    return [c for f6 in a for f2_ in f6 for c in f2_]
def f12_to_bytes(a): return b"".join(c.to_bytes(48, "little") for c in f12_coeffs(a))
def f12_from_bytes(b):
    c = [int.from_bytes(b[48 * i:48 * i + 48], "little") for i in range(12)]
    return (((c[0], c[1]), (c[2], c[3]), (c[4], c[5])), ((c[6], c[7]), (c[8], c[9]), (c[10], c[11])))


# ---- curves -----------------------------------------------------------------------------------------------------------------
def g1_add(P, S):
    if P is None: return S
    if S is None: return P
    if P[0] == S[0]:
        if (P[1] + S[1]) % Q == 0: return None
        l = 3 * P[0] * P[0] * pow(2 * P[1], -1, Q) % Q
    else:
        l = (S[1] - P[1]) * pow(S[0] - P[0], -1, Q) % Q
    x = (l * l - P[0] - S[0]) % Q
    return (x, (l * (P[0] - x) - P[1]) % Q)
def g1_neg(P): return None if P is None else (P[0], (-P[1]) % Q)
def g1_mul(k, P):
    k %= R; acc = None
    while k:
        if k & 1: acc = g1_add(acc, P)
        P = g1_add(P, P); k >>= 1
    return acc
def g1_on_curve(P): return P is None or (P[1] * P[1] - P[0] ** 3 - 4) % Q == 0

B2 = f2_scale(XI, 4)                              # the twist E': y^2 = x^3 + 4 (u + 1)
def g2_add(P, S):
    if P is None: return S
    if S is None: return P
    if P[0] == S[0]:
        if f2_add(P[1], S[1]) == F2_ZERO: return None
        l = f2_mul(f2_scale(f2_sqr(P[0]), 3), f2_inv(f2_scale(P[1], 2)))
    else:
        l = f2_mul(f2_sub(S[1], P[1]), f2_inv(f2_sub(S[0], P[0])))
    x = f2_sub(f2_sub(f2_sqr(l), P[0]), S[0])
    return (x, f2_sub(f2_mul(l, f2_sub(P[0], x)), P[1]))
def g2_neg(P): return None if P is None else (P[0], f2_neg(P[1]))
def g2_mul(k, P):
    k %= R; acc = None
    while k:
        if k & 1: acc = g2_add(acc, P)
        P = g2_add(P, P); k >>= 1
    return acc
def g2_on_curve(P): return P is None or f2_sub(f2_sqr(P[1]), f2_add(f2_mul(f2_sqr(P[0]), P[0]), B2)) == F2_ZERO

def g1_to_bytes(P): return bytes(96) if P is None else P[0].to_bytes(48, "little") + P[1].to_bytes(48, "little")
def g1_from_bytes(b):
    b = bytes(b)
    return None if not any(b) else (int.from_bytes(b[:48], "little"), int.from_bytes(b[48:96], "little"))
def g2_to_bytes(P): return bytes(192) if P is None else b"".join(c.to_bytes(48, "little") for c in (P[0][0], P[0][1], P[1][0], P[1][1]))
def g2_from_bytes(b):
    b = bytes(b)
    if not any(b): return None
    c = [int.from_bytes(b[48 * i:48 * i + 48], "little") for i in range(4)]
    return ((c[0], c[1]), (c[2], c[3]))


# ---- pairing ----------------------------------------------------------------------------------------------------------------
def _w_powers(c0, c2, c3):
    """the F_q12 element c0 + c2 w^2 + c3 w^3 with c0, c2, c3 in F_q2 (w^2 = v, w^3 = v w)"""
    return ((c0, c2, F2_ZERO), (F2_ZERO, c3, F2_ZERO))

def _line(T, S, P):
    """line through the untwisted images of T, S in E'(F_q2) (tangent if T == S), evaluated at P in E(F_q), scaled by w^3
    (an element of a proper subfield, removed by the final exponentiation): with (x, y) -> (x / w^2, y / w^3),
    l * w^3 = (lambda x_T - y_T) - lambda x_P w^2 + y_P w^3."""
    if T[0] == S[0] and f2_add(T[1], S[1]) == F2_ZERO:           # vertical: x_P - x_T / w^2, scaled by w^2
        return _w_powers(f2_neg(T[0]), (P[0], 0), F2_ZERO)
    if T == S:
        lam = f2_mul(f2_scale(f2_sqr(T[0]), 3), f2_inv(f2_scale(T[1], 2)))
    else:
        lam = f2_mul(f2_sub(S[1], T[1]), f2_inv(f2_sub(S[0], T[0])))
    return _w_powers(f2_sub(f2_mul(lam, T[0]), T[1]), f2_neg(f2_scale(lam, P[0])), (P[1], 0))

def miller_loop(P, Qp):
    """f_{|x|, Q}(P), conjugated because x < 0; 1 if either point is the identity"""
    if P is None or Qp is None:
        return F12_ONE
    f, T = F12_ONE, Qp
    for bit in bin(X_ABS)[3:]:
        f = f12_mul(f12_sqr(f), _line(T, T, P)); T = g2_add(T, T)
        if bit == "1":
            f = f12_mul(f, _line(T, Qp, P)); T = g2_add(T, Qp)
    return f12_conj(f)

FINAL_EXP = (Q ** 12 - 1) // R
def final_exponentiation(f): return f12_pow(f, FINAL_EXP)
def pairing(P, Qp): return final_exponentiation(miller_loop(P, Qp))
def gt_cubed(e): return f12_mul(f12_sqr(e), e)
def pairing_product_is_one(pairs):
    f = F12_ONE
    for P, Qp in pairs:
        f = f12_mul(f, miller_loop(P, Qp))
    return final_exponentiation(f) == F12_ONE


# ---- KZG (one opening per commitment) ----------------------------------------------------------------------------
def kzg_commit(coeffs, tau):
    """[p(tau)] G1 for a PUBLIC tau (test SRS)"""
    acc, t = 0, 1
    for c in coeffs:
        acc = (acc + c * t) % R; t = t * tau % R
    return g1_mul(acc, G1_GEN)

def kzg_open(coeffs, z, tau):
    """(v, W): v = p(z), W = [(p(tau) - v) / (tau - z)] G1"""
    v = 0
    for c in reversed(coeffs):
        v = (v * z + c) % R
    ptau, t = 0, 1
    for c in coeffs:
        ptau = (ptau + c * t) % R; t = t * tau % R
    return v, g1_mul((ptau - v) * pow(tau - z, -1, R) % R, G1_GEN)

def kzg_batch_verify(commitments, zs, vs, proofs, rs, g2, tau_g2):
    """sum_i r_i (C_i - [v_i] G1 + [z_i] W_i) paired with G2 against sum_i r_i W_i paired with [tau] G2: the aggregated form of
    e(C_i - [v_i] G1, G2) = e(W_i, [tau - z_i] G2) with caller-supplied coefficients r_i (the transcript's challenges)"""
    L, Rr = None, None
    sv = 0
    for C, z, v, W, r in zip(commitments, zs, vs, proofs, rs):
        L = g1_add(L, g1_add(g1_mul(r, C), g1_mul(r * z % R, W)))
        Rr = g1_add(Rr, g1_mul(r, W))
        sv = (sv + r * v) % R
    L = g1_add(L, g1_mul((-sv) % R, G1_GEN))
    return pairing_product_is_one([(L, g2), (g1_neg(Rr), tau_g2)])
