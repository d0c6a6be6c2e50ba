// the BLS12-381 pairing and the batched KZG opening check behind `ring` -> ring-proof's verifier
// (`Bls12::multi_miller_loop` + `final_exponentiation` of ark-ec / ark-bls12-381, named through the re-exports at
// /root/reference/src/lib.rs:13-17: `ring`, `ring_suite_types`).  The pairing VALUE is fixed by mathematics up to the choice
// of the final exponent; the engine's hard part carries the usual factor 3 (its GT values are the cube of the textbook
// e(P, Q)^((q^12-1)/r)), which no equality check between products of pairings can see.
//
// Tower: F_q2 = F_q[u]/(u^2+1), F_q6 = F_q2[v]/(v^3 - (u+1)), F_q12 = F_q6[w]/(w^2 - v) on the 12-limb Montgomery F_q of
// arith.cuh.  Miller loop: homogeneous projective doubling / addition steps on the M-type twist y^2 = x^3 + 4(u+1) with the
// line folded in as a sparse product (coefficients at w^0, w^2 (x x_P), w^3 (x y_P) -> positions 0, 1, 4 of the tower).
// Final exponentiation: easy part by one inversion + Frobenius, hard part by the chain
//     3 (q^4 - q^2 + 1)/r = (x - 1)^2 (x + q)(x^2 + q^2 - 1) + 3      (x = -0xd201000000010000)
// with Granger-Scott squarings in the cyclotomic subgroup.  Every formula here has a line-by-line Python twin in
// tools/gen_pairing_consts.py that is checked against the naive oracle (oracle/pairing_ref.py) before the constants are emitted.
//
// This header is the one-thread-per-pairing statement of the arithmetic (also compiled for the host by tests/host_emul); the
// warp-cooperative kernels of pairing_coop.cuh run the same formulas with the F_q2 products of every step spread over lanes.
#pragma once
#include "msm.cuh"
#include "gen/pairing_consts.cuh"

namespace vrfs {

struct Fq2 { Fq381 c0, c1; };
struct Fq6 { Fq2 c0, c1, c2; };
struct Fq12 { Fq6 c0, c1; };

template <uint32_t (*Fn)(int)> HD_INLINE Fq381 pconst() { Fq381 r; for (int i = 0; i < 12; i++) r.v[i] = Fn(i); return r; }
#define PAIRING_F2_CONST(NAME) (Fq2{pconst<PairingConsts::NAME##_C0>(), pconst<PairingConsts::NAME##_C1>()})

// ---- F_q2 -----------------------------------------------------------------------------------------------------------------
HD_INLINE Fq2 f2_zero() { return Fq2{Fq381::zero(), Fq381::zero()}; }
HD_INLINE Fq2 f2_one() { return Fq2{Fq381::one(), Fq381::zero()}; }
HD_INLINE bool f2_is_zero(const Fq2& a) { return a.c0.is_zero() & a.c1.is_zero(); }
HD_INLINE bool f2_eq(const Fq2& a, const Fq2& b) { return (a.c0 == b.c0) & (a.c1 == b.c1); }
HD_INLINE Fq2 f2_add(const Fq2& a, const Fq2& b) { return Fq2{a.c0 + b.c0, a.c1 + b.c1}; }
HD_INLINE Fq2 f2_sub(const Fq2& a, const Fq2& b) { return Fq2{a.c0 - b.c0, a.c1 - b.c1}; }
HD_INLINE Fq2 f2_neg(const Fq2& a) { return Fq2{neg(a.c0), neg(a.c1)}; }
HD_INLINE Fq2 f2_dbl(const Fq2& a) { return Fq2{dbl(a.c0), dbl(a.c1)}; }
HD_INLINE Fq2 f2_conj(const Fq2& a) { return Fq2{a.c0, neg(a.c1)}; }
HD_INLINE Fq2 f2_mul_xi(const Fq2& a) { return Fq2{a.c0 - a.c1, a.c0 + a.c1}; }          // * (u + 1)
HD_INLINE Fq2 f2_scale(const Fq2& a, const Fq381& k) { return Fq2{a.c0 * k, a.c1 * k}; }
HD_NOINLINE Fq2 f2_mul(const Fq2& a, const Fq2& b) {                                       // Karatsuba: 3 products
  Fq381 t0 = a.c0 * b.c0, t1 = a.c1 * b.c1, t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
  return Fq2{t0 - t1, t2 - t0 - t1};
}
HD_NOINLINE Fq2 f2_sqr(const Fq2& a) {                                                     // complex squaring: 2 products
  Fq381 t = a.c0 * a.c1;
  return Fq2{(a.c0 + a.c1) * (a.c0 - a.c1), dbl(t)};
}
HD_NOINLINE Fq2 f2_inv(const Fq2& a) {                                                     // 0 -> 0
  Fq381 d = fq381_inv_fast(sqr(a.c0) + sqr(a.c1));
  return Fq2{a.c0 * d, neg(a.c1 * d)};
}
HD_INLINE Fq2 f2_triple(const Fq2& a) { return f2_add(f2_dbl(a), a); }

// ---- F_q6 -----------------------------------------------------------------------------------------------------------------
HD_INLINE Fq6 f6_zero() { return Fq6{f2_zero(), f2_zero(), f2_zero()}; }
HD_INLINE Fq6 f6_one() { return Fq6{f2_one(), f2_zero(), f2_zero()}; }
HD_INLINE Fq6 f6_add(const Fq6& a, const Fq6& b) { return Fq6{f2_add(a.c0, b.c0), f2_add(a.c1, b.c1), f2_add(a.c2, b.c2)}; }
HD_INLINE Fq6 f6_sub(const Fq6& a, const Fq6& b) { return Fq6{f2_sub(a.c0, b.c0), f2_sub(a.c1, b.c1), f2_sub(a.c2, b.c2)}; }
HD_INLINE Fq6 f6_neg(const Fq6& a) { return Fq6{f2_neg(a.c0), f2_neg(a.c1), f2_neg(a.c2)}; }
HD_INLINE Fq6 f6_mul_v(const Fq6& a) { return Fq6{f2_mul_xi(a.c2), a.c0, a.c1}; }
HD_NOINLINE Fq6 f6_mul(const Fq6& a, const Fq6& b) {                                       // Karatsuba: 6 F_q2 products
  Fq2 v0 = f2_mul(a.c0, b.c0), v1 = f2_mul(a.c1, b.c1), v2 = f2_mul(a.c2, b.c2);
  Fq2 t0 = f2_sub(f2_sub(f2_mul(f2_add(a.c1, a.c2), f2_add(b.c1, b.c2)), v1), v2);        // a1 b2 + a2 b1
  Fq2 t1 = f2_sub(f2_sub(f2_mul(f2_add(a.c0, a.c1), f2_add(b.c0, b.c1)), v0), v1);        // a0 b1 + a1 b0
  Fq2 t2 = f2_sub(f2_sub(f2_mul(f2_add(a.c0, a.c2), f2_add(b.c0, b.c2)), v0), v2);        // a0 b2 + a2 b0
  return Fq6{f2_add(v0, f2_mul_xi(t0)), f2_add(t1, f2_mul_xi(v2)), f2_add(t2, v1)};
}
HD_NOINLINE Fq6 f6_mul_by_01(const Fq6& s, const Fq2& c0, const Fq2& c1) {                 // * (c0 + c1 v): 5 products
  Fq2 a_a = f2_mul(s.c0, c0), b_b = f2_mul(s.c1, c1);
  Fq2 t1 = f2_add(f2_mul_xi(f2_sub(f2_mul(f2_add(s.c1, s.c2), c1), b_b)), a_a);
  Fq2 t3 = f2_add(f2_sub(f2_mul(f2_add(s.c0, s.c2), c0), a_a), b_b);
  Fq2 t2 = f2_sub(f2_sub(f2_mul(f2_add(s.c0, s.c1), f2_add(c0, c1)), a_a), b_b);
  return Fq6{t1, t2, t3};
}
HD_NOINLINE Fq6 f6_mul_by_1(const Fq6& s, const Fq2& c1) {                                 // * (c1 v): 3 products
  return Fq6{f2_mul_xi(f2_mul(s.c2, c1)), f2_mul(s.c0, c1), f2_mul(s.c1, c1)};
}
HD_NOINLINE Fq6 f6_inv(const Fq6& a) {
  Fq2 t0 = f2_sub(f2_sqr(a.c0), f2_mul_xi(f2_mul(a.c1, a.c2)));
  Fq2 t1 = f2_sub(f2_mul_xi(f2_sqr(a.c2)), f2_mul(a.c0, a.c1));
  Fq2 t2 = f2_sub(f2_sqr(a.c1), f2_mul(a.c0, a.c2));
  Fq2 d = f2_inv(f2_add(f2_mul(a.c0, t0), f2_mul_xi(f2_add(f2_mul(a.c2, t1), f2_mul(a.c1, t2)))));
  return Fq6{f2_mul(t0, d), f2_mul(t1, d), f2_mul(t2, d)};
}
// a^(q^K), K = 1 or 2
template <int K> HD_INLINE Fq2 f2_frob(const Fq2& a) { return (K & 1) ? f2_conj(a) : a; }
template <int K> HD_INLINE Fq6 f6_frob(const Fq6& a) {
  const Fq2 g1 = K == 1 ? PAIRING_F2_CONST(G6_1_1) : PAIRING_F2_CONST(G6_1_2), g2 = K == 1 ? PAIRING_F2_CONST(G6_2_1) : PAIRING_F2_CONST(G6_2_2);
  return Fq6{f2_frob<K>(a.c0), f2_mul(f2_frob<K>(a.c1), g1), f2_mul(f2_frob<K>(a.c2), g2)};
}

// ---- F_q12 ----------------------------------------------------------------------------------------------------------------
HD_INLINE Fq12 f12_one() { return Fq12{f6_one(), f6_zero()}; }
HD_INLINE Fq12 f12_conj(const Fq12& a) { return Fq12{a.c0, f6_neg(a.c1)}; }
HD_INLINE bool f12_is_one(const Fq12& a) {
  const Fq381 one = Fq381::one();
  bool ok = (a.c0.c0.c0 == one) & a.c0.c0.c1.is_zero();
  ok &= f2_is_zero(a.c0.c1) & f2_is_zero(a.c0.c2) & f2_is_zero(a.c1.c0) & f2_is_zero(a.c1.c1) & f2_is_zero(a.c1.c2);
  return ok;
}
HD_NOINLINE Fq12 f12_mul(const Fq12& a, const Fq12& b) {                                   // Karatsuba: 3 F_q6 products
  Fq6 aa = f6_mul(a.c0, b.c0), bb = f6_mul(a.c1, b.c1);
  Fq6 c1 = f6_sub(f6_sub(f6_mul(f6_add(a.c0, a.c1), f6_add(b.c0, b.c1)), aa), bb);
  return Fq12{f6_add(aa, f6_mul_v(bb)), c1};
}
HD_NOINLINE Fq12 f12_sqr(const Fq12& a) {                                                  // complex squaring: 2 F_q6 products
  Fq6 ab = f6_mul(a.c0, a.c1);
  Fq6 t = f6_mul(f6_add(a.c0, a.c1), f6_add(a.c0, f6_mul_v(a.c1)));
  return Fq12{f6_sub(f6_sub(t, ab), f6_mul_v(ab)), f6_add(ab, ab)};
}
HD_NOINLINE Fq12 f12_inv(const Fq12& a) {
  Fq6 d = f6_inv(f6_sub(f6_mul(a.c0, a.c0), f6_mul_v(f6_mul(a.c1, a.c1))));
  return Fq12{f6_mul(a.c0, d), f6_neg(f6_mul(a.c1, d))};
}
template <int K> HD_NOINLINE Fq12 f12_frob(const Fq12& a) {
  const Fq2 g = K == 1 ? PAIRING_F2_CONST(G12_1) : PAIRING_F2_CONST(G12_2);
  Fq6 c1 = f6_frob<K>(a.c1);
  return Fq12{f6_frob<K>(a.c0), Fq6{f2_mul(c1.c0, g), f2_mul(c1.c1, g), f2_mul(c1.c2, g)}};
}
// f * (c0 + c1 v + c4 v w): the line of a Miller step (13 F_q2 products)
HD_NOINLINE Fq12 f12_mul_by_014(const Fq12& f, const Fq2& c0, const Fq2& c1, const Fq2& c4) {
  Fq6 aa = f6_mul_by_01(f.c0, c0, c1), bb = f6_mul_by_1(f.c1, c4);
  Fq6 n1 = f6_sub(f6_sub(f6_mul_by_01(f6_add(f.c1, f.c0), c0, f2_add(c1, c4)), aa), bb);
  return Fq12{f6_add(f6_mul_v(bb), aa), n1};
}
// Granger-Scott squaring, valid in the cyclotomic subgroup (after the easy part): 9 F_q2 squarings
HD_INLINE void fp4_sqr(Fq2& r0, Fq2& r1, const Fq2& x, const Fq2& y) {
  Fq2 t0 = f2_sqr(x), t1 = f2_sqr(y);
  r0 = f2_add(f2_mul_xi(t1), t0);
  r1 = f2_sub(f2_sub(f2_sqr(f2_add(x, y)), t0), t1);
}
HD_NOINLINE Fq12 f12_cyclotomic_sqr(const Fq12& a) {
  Fq2 z0 = a.c0.c0, z4 = a.c0.c1, z3 = a.c0.c2, z2 = a.c1.c0, z1 = a.c1.c1, z5 = a.c1.c2, t0, t1, t2, t3;
  fp4_sqr(t0, t1, z0, z1);
  z0 = f2_add(f2_dbl(f2_sub(t0, z0)), t0);
  z1 = f2_add(f2_dbl(f2_add(t1, z1)), t1);
  fp4_sqr(t0, t1, z2, z3);
  fp4_sqr(t2, t3, z4, z5);
  z4 = f2_add(f2_dbl(f2_sub(t0, z4)), t0);
  z5 = f2_add(f2_dbl(f2_add(t1, z5)), t1);
  t0 = f2_mul_xi(t3);
  z2 = f2_add(f2_dbl(f2_add(t0, z2)), t0);
  z3 = f2_add(f2_dbl(f2_sub(t2, z3)), t2);
  return Fq12{Fq6{z0, z4, z3}, Fq6{z2, z1, z5}};
}

// ---- G2 on the twist, Miller loop ------------------------------------------------------------------------------------------
struct G2Aff { Fq2 x, y; bool inf; };
struct G2Proj { Fq2 X, Y, Z; };
struct G1AffPt { Fq381 x, y; bool inf; };
struct LineCoeffs { Fq2 c0, c1, c4; };             // c1 is scaled by x_P, c4 by y_P before the sparse product

HD_INLINE Fq2 g2_b() { Fq381 four = dbl(dbl(Fq381::one())); return Fq2{four, four}; }       // 4 (u + 1)
HD_INLINE bool g2_on_curve(const G2Aff& p) {
  if (p.inf) return true;
  return f2_eq(f2_sqr(p.y), f2_add(f2_mul(f2_sqr(p.x), p.x), g2_b()));
}
// 192 bytes x.c0 | x.c1 | y.c0 | y.c1 (48-byte LE canonical; zeros = identity); false if a coordinate is >= q
HD_INLINE bool g2_load(G2Aff& p, const uint8_t* b) {
  uint32_t raw[4][12];
  bool ok = true;
  uint32_t any = 0;
  for (int k = 0; k < 4; k++) { load_le<12>(raw[k], b + 48 * k); ok &= is_canonical<BlsFq>(raw[k]); for (int i = 0; i < 12; i++) any |= raw[k][i]; }
  p.x = Fq2{to_mont<BlsFq>(raw[0]), to_mont<BlsFq>(raw[1])};
  p.y = Fq2{to_mont<BlsFq>(raw[2]), to_mont<BlsFq>(raw[3])};
  p.inf = any == 0;
  return ok;
}
HD_INLINE bool g1_load_bytes(G1AffPt& p, const uint8_t* b) {
  uint32_t rx[12], ry[12];
  load_le<12>(rx, b); load_le<12>(ry, b + 48);
  uint32_t any = 0;
  for (int i = 0; i < 12; i++) any |= rx[i] | ry[i];
  bool ok = is_canonical<BlsFq>(rx) & is_canonical<BlsFq>(ry);
  p.x = to_mont<BlsFq>(rx); p.y = to_mont<BlsFq>(ry); p.inf = any == 0;
  return ok;
}
HD_INLINE bool g1_on_curve_pt(const G1AffPt& p) { return p.inf || sw_on_curve<G1Curve>(p.x, p.y); }

HD_NOINLINE void g2_doubling_step(G2Proj& r, LineCoeffs& l) {
  const Fq381 two_inv = pconst<PairingConsts::TWO_INV>();
  Fq2 a = f2_scale(f2_mul(r.X, r.Y), two_inv);
  Fq2 b = f2_sqr(r.Y), c = f2_sqr(r.Z);
  Fq2 e = f2_mul(g2_b(), f2_triple(c));
  Fq2 f = f2_triple(e);
  Fq2 g = f2_scale(f2_add(b, f), two_inv);
  Fq2 h = f2_sub(f2_sqr(f2_add(r.Y, r.Z)), f2_add(b, c));
  Fq2 i = f2_sub(e, b);
  Fq2 j = f2_sqr(r.X);
  Fq2 e2 = f2_sqr(e);
  r.X = f2_mul(a, f2_sub(b, f));
  r.Y = f2_sub(f2_sqr(g), f2_triple(e2));
  r.Z = f2_mul(b, h);
  l.c0 = i; l.c1 = f2_triple(j); l.c4 = f2_neg(h);
}
HD_NOINLINE void g2_addition_step(G2Proj& r, const G2Aff& q, LineCoeffs& l) {
  Fq2 theta = f2_sub(r.Y, f2_mul(q.y, r.Z));
  Fq2 lam = f2_sub(r.X, f2_mul(q.x, r.Z));
  Fq2 c = f2_sqr(theta), d = f2_sqr(lam), e = f2_mul(lam, d), f = f2_mul(r.Z, c), g = f2_mul(r.X, d);
  Fq2 h = f2_sub(f2_add(e, f), f2_dbl(g));
  Fq2 Y3 = f2_sub(f2_mul(theta, f2_sub(g, h)), f2_mul(e, r.Y));
  r.X = f2_mul(lam, h);
  r.Y = Y3;
  r.Z = f2_mul(r.Z, e);
  l.c0 = f2_sub(f2_mul(theta, q.x), f2_mul(lam, q.y)); l.c1 = f2_neg(theta); l.c4 = lam;
}
HD_INLINE Fq12 ell(const Fq12& f, const LineCoeffs& l, const G1AffPt& p) {
  return f12_mul_by_014(f, l.c0, f2_scale(l.c1, p.x), f2_scale(l.c4, p.y));
}
#define PAIRING_MAX_PAIRS 4
// prod_k f_{|x|, Q_k}(P_k), conjugated (x < 0); pairs with an identity contribute 1
HD_NOINLINE Fq12 multi_miller_loop(int n, const G1AffPt* ps, const G2Aff* qs) {
  G2Proj r[PAIRING_MAX_PAIRS];
  bool live[PAIRING_MAX_PAIRS];
  for (int k = 0; k < n; k++) { live[k] = !(ps[k].inf | qs[k].inf); r[k] = G2Proj{qs[k].x, qs[k].y, f2_one()}; }
  Fq12 f = f12_one();
  LineCoeffs l;
#pragma unroll 1
  for (int bit = 62; bit >= 0; bit--) {
    f = f12_sqr(f);
    for (int k = 0; k < n; k++) if (live[k]) { g2_doubling_step(r[k], l); f = ell(f, l, ps[k]); }
    if ((PairingConsts::X_ABS >> bit) & 1ull)
      for (int k = 0; k < n; k++) if (live[k]) { g2_addition_step(r[k], qs[k], l); f = ell(f, l, ps[k]); }
  }
  return f12_conj(f);
}
// m^x for the (negative) curve parameter, m in the cyclotomic subgroup
HD_NOINLINE Fq12 f12_exp_by_x(const Fq12& m) {
  Fq12 r = m;
#pragma unroll 1
  for (int bit = 62; bit >= 0; bit--) {
    r = f12_cyclotomic_sqr(r);
    if ((PairingConsts::X_ABS >> bit) & 1ull) r = f12_mul(r, m);
  }
  return f12_conj(r);
}
HD_NOINLINE Fq12 final_exponentiation(const Fq12& f) {
  Fq12 r = f12_mul(f12_conj(f), f12_inv(f));                        // ^(q^6 - 1)
  Fq12 m = f12_mul(f12_frob<2>(r), r);                              // ^(q^2 + 1)
  Fq12 a = f12_mul(f12_exp_by_x(m), f12_conj(m));                   // ^(x - 1)
  Fq12 b = f12_mul(f12_exp_by_x(a), f12_conj(a));                   // ^(x - 1)
  Fq12 c = f12_mul(f12_exp_by_x(b), f12_frob<1>(b));                // ^(x + q)
  Fq12 d = f12_mul(f12_mul(f12_exp_by_x(f12_exp_by_x(c)), f12_frob<2>(c)), f12_conj(c));   // ^(x^2 + q^2 - 1)
  return f12_mul(d, f12_mul(f12_cyclotomic_sqr(m), m));             // * m^3
}
// 12 F_q coefficients, canonical 48-byte LE each (c0.c0.c0, c0.c0.c1, c0.c1.c0, ...)
HD_INLINE void f12_store(uint8_t* out, const Fq12& a) {
  const Fq381* c[12] = {&a.c0.c0.c0, &a.c0.c0.c1, &a.c0.c1.c0, &a.c0.c1.c1, &a.c0.c2.c0, &a.c0.c2.c1,
                        &a.c1.c0.c0, &a.c1.c0.c1, &a.c1.c1.c0, &a.c1.c1.c1, &a.c1.c2.c0, &a.c1.c2.c1};
  uint32_t raw[12];
  for (int k = 0; k < 12; k++) { from_mont<BlsFq>(raw, *c[k]); store_le<12>(out + 48 * k, raw); }
}
// one pairing check on one thread: prod e(P_k, Q_k) == 1, inputs as ABI bytes; `negate` bit k negates P_k.
// returns 1 accepted, 0 rejected, 2 malformed input (non-canonical coordinate / not on the curve)
HD_INLINE int pairing_product_check_bytes(int n, const uint8_t* g1, const uint8_t* g2, unsigned negate, Fq12* value_out = nullptr) {
  G1AffPt ps[PAIRING_MAX_PAIRS]; G2Aff qs[PAIRING_MAX_PAIRS];
  bool ok = true;
  for (int k = 0; k < n; k++) {
    ok &= g1_load_bytes(ps[k], g1 + 96 * k) & g2_load(qs[k], g2 + 192 * k);
    ok &= g1_on_curve_pt(ps[k]) & g2_on_curve(qs[k]);
    if ((negate >> k) & 1u) ps[k].y = neg(ps[k].y);
  }
  if (!ok) return 2;
  Fq12 e = final_exponentiation(multi_miller_loop(n, ps, qs));
  if (value_out) *value_out = e;
  return f12_is_one(e) ? 1 : 0;
}

}  // namespace vrfs
