#!/usr/bin/env python3
"""Emit ark_ec_vrfs_b200/csrc/gen/pairing_consts.cuh (Frobenius coefficients of the F_q12 tower, the twist constant, G2 generator)
and, before writing anything, CHECK the engine's pairing algorithm against the naive oracle (oracle/pairing_ref.py):
this file contains a line-by-line Python twin of csrc/pairing.cuh (projective doubling / addition steps with sparse line
products, Frobenius maps from the emitted constants, cyclotomic exponentiation by the curve parameter, the x-chain of the hard
part) and asserts that it gives gt_cubed(oracle pairing) on random inputs.  Test / build infrastructure only."""
import os
import random
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import pairing_ref as P  # noqa: E402

OUT = os.path.join(ROOT, "ark_ec_vrfs_b200", "csrc", "gen", "pairing_consts.cuh")
Q, R = P.Q, P.R

# Frobenius coefficients: gamma6_1[k] = xi^((q^k - 1)/3), gamma6_2[k] = xi^(2 (q^k - 1)/3), gamma12[k] = xi^((q^k - 1)/6), k = 1, 2
G6_1 = {k: P.f2_pow(P.XI, (Q ** k - 1) // 3) for k in (1, 2)}
G6_2 = {k: P.f2_pow(P.XI, 2 * (Q ** k - 1) // 3) for k in (1, 2)}
G12 = {k: P.f2_pow(P.XI, (Q ** k - 1) // 6) for k in (1, 2)}


# ---- twin of csrc/pairing.cuh ---------------------------------------------------------------------------------------------
def f2_frob(a, k): return P.f2_conj(a) if k & 1 else a
def f6_frob(a, k): return (f2_frob(a[0], k), P.f2_mul(f2_frob(a[1], k), G6_1[k]), P.f2_mul(f2_frob(a[2], k), G6_2[k]))
def f12_frob(a, k):
    c1 = f6_frob(a[1], k)
    return (f6_frob(a[0], k), tuple(P.f2_mul(c, G12[k]) for c in c1))

def f6_mul_by_01(s, c0, c1):
    a_a, b_b = P.f2_mul(s[0], c0), P.f2_mul(s[1], c1)
    t1 = P.f2_add(P.f2_mul_xi(P.f2_sub(P.f2_mul(P.f2_add(s[1], s[2]), c1), b_b)), a_a)
    t3 = P.f2_add(P.f2_sub(P.f2_mul(P.f2_add(s[0], s[2]), c0), a_a), b_b)
    t2 = P.f2_sub(P.f2_sub(P.f2_mul(P.f2_add(s[0], s[1]), P.f2_add(c0, c1)), a_a), b_b)
    return (t1, t2, t3)
def f6_mul_by_1(s, c1): return (P.f2_mul_xi(P.f2_mul(s[2], c1)), P.f2_mul(s[0], c1), P.f2_mul(s[1], c1))
def f12_mul_by_014(f, c0, c1, c4):
    aa = f6_mul_by_01(f[0], c0, c1)
    bb = f6_mul_by_1(f[1], c4)
    o = P.f2_add(c1, c4)
    n1 = P.f6_sub(P.f6_sub(f6_mul_by_01(P.f6_add(f[1], f[0]), c0, o), aa), bb)
    return (P.f6_add(P.f6_mul_v(bb), aa), n1)

TWO_INV = pow(2, -1, Q)
def doubling_step(r):
    """r = (X, Y, Z) homogeneous projective on the twist; returns (2r, line coefficients (c0, c1 [x x_P], c4 [x y_P]))"""
    X, Y, Z = r
    a = P.f2_scale(P.f2_mul(X, Y), TWO_INV)
    b = P.f2_sqr(Y); c = P.f2_sqr(Z)
    e = P.f2_mul(P.B2, P.f2_scale(c, 3))
    f = P.f2_scale(e, 3)
    g = P.f2_scale(P.f2_add(b, f), TWO_INV)
    h = P.f2_sub(P.f2_sqr(P.f2_add(Y, Z)), P.f2_add(b, c))
    i = P.f2_sub(e, b)
    j = P.f2_sqr(X)
    e2 = P.f2_sqr(e)
    X3 = P.f2_mul(a, P.f2_sub(b, f))
    Y3 = P.f2_sub(P.f2_sqr(g), P.f2_scale(e2, 3))
    Z3 = P.f2_mul(b, h)
    return (X3, Y3, Z3), (i, P.f2_scale(j, 3), P.f2_neg(h))
def addition_step(r, q):
    X, Y, Z = r
    theta = P.f2_sub(Y, P.f2_mul(q[1], Z))
    lam = P.f2_sub(X, P.f2_mul(q[0], Z))
    c = P.f2_sqr(theta); d = P.f2_sqr(lam); e = P.f2_mul(lam, d); f = P.f2_mul(Z, c); g = P.f2_mul(X, d)
    h = P.f2_sub(P.f2_add(e, f), P.f2_scale(g, 2))
    X3 = P.f2_mul(lam, h)
    Y3 = P.f2_sub(P.f2_mul(theta, P.f2_sub(g, h)), P.f2_mul(e, Y))
    Z3 = P.f2_mul(Z, e)
    j = P.f2_sub(P.f2_mul(theta, q[0]), P.f2_mul(lam, q[1]))
    return (X3, Y3, Z3), (j, P.f2_neg(theta), lam)
def ell(f, coeffs, p):
    return f12_mul_by_014(f, coeffs[0], P.f2_scale(coeffs[1], p[0]), P.f2_scale(coeffs[2], p[1]))
def multi_miller(pairs):
    pairs = [(p, q) for p, q in pairs if p is not None and q is not None]
    f = P.F12_ONE
    rs = [(q[0], q[1], P.F2_ONE) for _, q in pairs]
    for bit in bin(P.X_ABS)[3:]:
        f = P.f12_sqr(f)
        for k, (p, q) in enumerate(pairs):
            rs[k], co = doubling_step(rs[k]); f = ell(f, co, p)
        if bit == "1":
            for k, (p, q) in enumerate(pairs):
                rs[k], co = addition_step(rs[k], q); f = ell(f, co, p)
    return P.f12_conj(f)
def exp_by_x(m):
    """m^x for x = -X_ABS, m in the cyclotomic subgroup (inverse = conjugate)"""
    r = m
    for bit in bin(P.X_ABS)[3:]:
        r = P.f12_sqr(r)
        if bit == "1":
            r = P.f12_mul(r, m)
    return P.f12_conj(r)
def final_exp(f):
    r = P.f12_mul(P.f12_conj(f), P.f12_inv(f))           # f^(q^6 - 1)
    m = P.f12_mul(f12_frob(r, 2), r)                     # ^(q^2 + 1)
    a = P.f12_mul(exp_by_x(m), P.f12_conj(m))            # m^(x - 1)
    b = P.f12_mul(exp_by_x(a), P.f12_conj(a))            # ^(x - 1)
    c = P.f12_mul(exp_by_x(b), f12_frob(b, 1))           # ^(x + q)
    d = P.f12_mul(P.f12_mul(exp_by_x(exp_by_x(c)), f12_frob(c, 2)), P.f12_conj(c))     # ^(x^2 + q^2 - 1)
    return P.f12_mul(d, P.f12_mul(P.f12_sqr(m), m))      # * m^3:   m^(3 (q^4 - q^2 + 1) / r)


def cyclotomic_sqr(a):
    """Granger-Scott squaring in the cyclotomic subgroup, on the tower (c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2) = (z0, z4, z3, z2, z1, z5)"""
    z0, z4, z3 = a[0]; z2, z1, z5 = a[1]
    def fp4_sqr(x, y):
        t0, t1 = P.f2_sqr(x), P.f2_sqr(y)
        return P.f2_add(P.f2_mul_xi(t1), t0), P.f2_sub(P.f2_sub(P.f2_sqr(P.f2_add(x, y)), t0), t1)
    t0, t1 = fp4_sqr(z0, z1)
    z0 = P.f2_add(P.f2_scale(P.f2_sub(t0, z0), 2), t0)
    z1 = P.f2_add(P.f2_scale(P.f2_add(t1, z1), 2), t1)
    t0, t1 = fp4_sqr(z2, z3)
    t2, t3 = fp4_sqr(z4, z5)
    z4 = P.f2_add(P.f2_scale(P.f2_sub(t0, z4), 2), t0)
    z5 = P.f2_add(P.f2_scale(P.f2_add(t1, z5), 2), t1)
    t0 = P.f2_mul_xi(t3)
    z2 = P.f2_add(P.f2_scale(P.f2_add(t0, z2), 2), t0)
    z3 = P.f2_add(P.f2_scale(P.f2_sub(t2, z3), 2), t2)
    return ((z0, z4, z3), (z2, z1, z5))


def self_check():
    rnd = random.Random(2026)
    rf2 = lambda: (rnd.randrange(Q), rnd.randrange(Q))
    rf6 = lambda: (rf2(), rf2(), rf2())
    a = (rf6(), rf6())
    for k in (1, 2):
        assert f12_frob(a, k) == P.f12_pow(a, Q ** k), "frobenius %d" % k
    c0, c1, c4 = rf2(), rf2(), rf2()
    assert f12_mul_by_014(a, c0, c1, c4) == P.f12_mul(a, ((c0, c1, P.F2_ZERO), (P.F2_ZERO, c4, P.F2_ZERO))), "mul_by_014"
    # projective steps against the affine group law
    q = P.g2_mul(rnd.randrange(R), P.G2_GEN)
    r2, _ = doubling_step((q[0], q[1], P.F2_ONE))
    zi = P.f2_inv(r2[2]); assert (P.f2_mul(r2[0], zi), P.f2_mul(r2[1], zi)) == P.g2_add(q, q), "doubling"
    r3, _ = addition_step(r2, q)
    zi = P.f2_inv(r3[2]); assert (P.f2_mul(r3[0], zi), P.f2_mul(r3[1], zi)) == P.g2_mul(3, q), "addition"
    for _ in range(2):
        s, t = rnd.randrange(1, R), rnd.randrange(1, R)
        p1, q1 = P.g1_mul(s, P.G1_GEN), P.g2_mul(t, P.G2_GEN)
        got = final_exp(multi_miller([(p1, q1)]))
        assert got == P.gt_cubed(P.pairing(p1, q1)), "pairing"
        m = P.f12_mul(f12_frob(P.f12_mul(P.f12_conj(a), P.f12_inv(a)), 2), P.f12_mul(P.f12_conj(a), P.f12_inv(a)))
        assert cyclotomic_sqr(m) == P.f12_sqr(m), "cyclotomic squaring"
    # product of pairings: e(aG1, G2) e(-G1, aG2) = 1
    s = rnd.randrange(1, R)
    assert final_exp(multi_miller([(P.g1_mul(s, P.G1_GEN), P.G2_GEN), (P.g1_neg(P.G1_GEN), P.g2_mul(s, P.G2_GEN))])) == P.F12_ONE
    print("pairing twin == oracle (Frobenius maps, sparse products, projective steps, x-chain, cyclotomic squaring)")


def limbs(x, n=12):
    return ", ".join("0x%08xu" % ((x >> (32 * i)) & 0xFFFFFFFF) for i in range(n))


def main():
    self_check()
    Rm = (1 << 384) % Q
    mont = lambda x: x * Rm % Q
    L = ["// generated by tools/gen_pairing_consts.py - do not edit", "#pragma once", '#include "field_consts.cuh"', "namespace vrfs {",
         "// Montgomery-form limbs (R = 2^384) of the constants of the F_q12 tower over BLS12-381 F_q; F_q2 values as c0 | c1",
         "struct PairingConsts {"]
    def acc(name, val):
        L.append(f"  static HD_INLINE uint32_t {name}(int i) {{ constexpr uint32_t t[12] = {{{limbs(mont(val))}}}; return t[i]; }}")
    for k in (1, 2):
        for nm, tab in (("G6_1", G6_1), ("G6_2", G6_2), ("G12", G12)):
            acc(f"{nm}_{k}_C0", tab[k][0]); acc(f"{nm}_{k}_C1", tab[k][1])
    acc("TWO_INV", TWO_INV)
    for nm, v in (("G2X_C0", P.G2_GEN[0][0]), ("G2X_C1", P.G2_GEN[0][1]), ("G2Y_C0", P.G2_GEN[1][0]), ("G2Y_C1", P.G2_GEN[1][1])):
        acc(nm, v)
    L.append(f"  static constexpr unsigned long long X_ABS = 0x{P.X_ABS:x}ull;   // the curve parameter is -X_ABS")
    L.append("};")
    L.append("}  // namespace vrfs")
    open(OUT, "w").write("\n".join(L) + "\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
