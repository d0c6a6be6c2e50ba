// K3 of SURVEY.md 2.5: twisted-Edwards points  a*x^2 + y^2 = 1 + d*x^2*y^2  in extended coordinates
// (X:Y:Z:T), T = XY/Z - the device replacement for ark-ec's `twisted_edwards::Projective` behind
// `AffinePoint<S>` (/root/reference/src/lib.rs:13-17) for Bandersnatch (a = -5) and Ed25519 (a = -1).
// Formulas: add-2008-hwcd / dbl-2008-hwcd (unified; exception-free on the prime-order subgroup, and on
// the whole curve when a is a square and d is not - Ed25519).
#pragma once
#include "gen/field_consts.cuh"

namespace vrfs {

// a generated constant (limb accessor) as a field element
template <class P, uint32_t (*Fn)(int)> HD_INLINE Fp<P> fconst() { Fp<P> r; for (int i = 0; i < P::N; i++) r.v[i] = Fn(i); return r; }

struct BandCurve {
  typedef BlsFr Fq;
  typedef BandFr Fr;
  typedef BandConsts K;
  typedef Fp<Fq> F;
  static constexpr int COF_LOG2 = 2;
  static constexpr bool IS_TE = true;
  static constexpr bool A_IS_M1 = false;
  static constexpr bool HAS_GLV = true;
  static HD_INLINE F mul_a(const F& x) { F t = dbl(dbl(x)); return neg(t + x); }   // a = -5
  static HD_INLINE F mul_neg_a(const F& x) { return dbl(dbl(x)) + x; }              // -a x = 5x (no negation: see te_dbl)
  static HD_INLINE F d() { return fconst<Fq, K::D>(); }
  static HD_INLINE F gx() { return fconst<Fq, K::GX>(); }
  static HD_INLINE F gy() { return fconst<Fq, K::GY>(); }
  static HD_INLINE F bx() { return fconst<Fq, K::BX>(); }
  static HD_INLINE F by() { return fconst<Fq, K::BY>(); }
};
struct EdCurve {
  typedef F25519 Fq;
  typedef EdFr Fr;
  typedef EdConsts K;
  typedef Fp<Fq> F;
  static constexpr int COF_LOG2 = 3;
  static constexpr bool IS_TE = true;
  static constexpr bool A_IS_M1 = true;
  static constexpr bool HAS_GLV = false;
  static HD_INLINE F mul_a(const F& x) { return neg(x); }                           // a = -1
  static HD_INLINE F mul_neg_a(const F& x) { return x; }
  static HD_INLINE F d() { return fconst<Fq, K::D>(); }
  static HD_INLINE F gx() { return fconst<Fq, K::GX>(); }
  static HD_INLINE F gy() { return fconst<Fq, K::GY>(); }
  static HD_INLINE F bx() { return fconst<Fq, K::BX>(); }
  static HD_INLINE F by() { return fconst<Fq, K::BY>(); }
};

// Jubjub (a = -1 over BLS12-381 Fr) and Baby-Jubjub (a = 1 over BN254 Fr); both have a square `a` and a non-square
// `d`, so the unified hwcd formulas are complete on the whole curve
struct JubCurve {
  typedef BlsFr Fq;
  typedef JubFr Fr;
  typedef JubConsts K;
  typedef Fp<Fq> F;
  static constexpr int COF_LOG2 = 3;
  static constexpr bool IS_TE = true;
  static constexpr bool A_IS_M1 = true;
  static constexpr bool HAS_GLV = false;
  static HD_INLINE F mul_a(const F& x) { return neg(x); }
  static HD_INLINE F mul_neg_a(const F& x) { return x; }
  static HD_INLINE F d() { return fconst<Fq, K::D>(); }
  static HD_INLINE F gx() { return fconst<Fq, K::GX>(); }
  static HD_INLINE F gy() { return fconst<Fq, K::GY>(); }
  static HD_INLINE F bx() { return fconst<Fq, K::BX>(); }
  static HD_INLINE F by() { return fconst<Fq, K::BY>(); }
};
struct BjjCurve {
  typedef Bn254Fr Fq;
  typedef BjjFr Fr;
  typedef BjjConsts K;
  typedef Fp<Fq> F;
  static constexpr int COF_LOG2 = 3;
  static constexpr bool IS_TE = true;
  static constexpr bool A_IS_M1 = false;
  static constexpr bool HAS_GLV = false;
  static HD_INLINE F mul_a(const F& x) { return x; }
  static HD_INLINE F mul_neg_a(const F& x) { return neg(x); }
  static HD_INLINE F d() { return fconst<Fq, K::D>(); }
  static HD_INLINE F gx() { return fconst<Fq, K::GX>(); }
  static HD_INLINE F gy() { return fconst<Fq, K::GY>(); }
  static HD_INLINE F bx() { return fconst<Fq, K::BX>(); }
  static HD_INLINE F by() { return fconst<Fq, K::BY>(); }
};

template <class C> struct TEPoint { typename C::F X, Y, Z, T; };
// table entry forms: "cached" = (X, Y, Z, d*T); "affine cached" = (x, y, d*x*y) with Z = 1.
// Curves with a = -1 (C::A_IS_M1: Ed25519, Jubjub) keep the same structs but store (Y+X, Y-X, 2Z, 2d*T) and (y+x, y-x, 2d*x*y):
// add-2008-hwcd-3 then takes 8 products (7 against an affine entry) instead of 9 (8), five field additions fewer, and negating an
// entry is a swap of its first two fields.
template <class C> struct TECached { typename C::F X, Y, Z, dT; };
template <class C> struct TEAffCached { typename C::F x, y, dt; };

template <class C> HD_INLINE void te_set_identity(TEPoint<C>& P) {
  P.X = C::F::zero(); P.Y = C::F::one(); P.Z = C::F::one(); P.T = C::F::zero();
}
template <class C> HD_INLINE void te_from_affine(TEPoint<C>& P, const typename C::F& x, const typename C::F& y) {
  P.X = x; P.Y = y; P.Z = C::F::one(); P.T = x * y;
}
template <class C> HD_INLINE bool te_is_identity(const TEPoint<C>& P) { return P.X.is_zero() && P.Y == P.Z; }
template <class C> HD_INLINE bool te_on_curve(const typename C::F& x, const typename C::F& y) {
  typename C::F xx = sqr(x), yy = sqr(y);
  return C::mul_a(xx) + yy == C::F::one() + C::d() * xx * yy;
}
template <class C> HD_INLINE void te_to_cached(TECached<C>& r, const TEPoint<C>& P) {
  if constexpr (C::A_IS_M1) { r.X = P.Y + P.X; r.Y = P.Y - P.X; r.Z = dbl(P.Z); r.dT = dbl(P.T * C::d()); }
  else { r.X = P.X; r.Y = P.Y; r.Z = P.Z; r.dT = P.T * C::d(); }
}
// the point a cached entry stands for, as (X : Y : Z) up to a common factor (what a doubling reads)
template <class C> HD_INLINE void te_from_cached_xyz(TEPoint<C>& P, const TECached<C>& e) {
  if constexpr (C::A_IS_M1) { P.X = e.X - e.Y; P.Y = e.X + e.Y; P.Z = e.Z; }      // (2X : 2Y : 2Z)
  else { P.X = e.X; P.Y = e.Y; P.Z = e.Z; }
  P.T = P.Z;                                                                        // not read by the doubling
}

// r = p + q (9M + 1 by d)
template <class C> HD_NOINLINE void te_add(TEPoint<C>* r, const TEPoint<C>* p, const TEPoint<C>* q) {
  typedef typename C::F F;
  F A = p->X * q->X, B = p->Y * q->Y, Cc = p->T * q->T * C::d(), D = p->Z * q->Z;
  F E = (p->X + p->Y) * (q->X + q->Y) - A - B;
  F Fv = D - Cc, G = D + Cc, H = B + C::mul_neg_a(A);
  r->X = E * Fv; r->Y = G * H; r->T = E * H; r->Z = Fv * G;
}
// r = p + q, q cached, optionally negated (9M)
// (want_t = false: T of the result is not computed - the next operation is a doubling, which does not read it, or the end)
template <class C> HD_NOINLINE void te_add_cached(TEPoint<C>* r, const TEPoint<C>* p, const TECached<C>* q, bool negate, bool want_t = true) {
  typedef typename C::F F;
  F A, B, Cc, D, E, Fv, G, H, X3, Y3, Z3;
  if constexpr (C::A_IS_M1) {
    F ypx = select(negate, q->Y, q->X), ymx = select(negate, q->X, q->Y), q2dT = cneg(q->dT, negate);
    mul2(A, B, p->Y - p->X, ymx, p->Y + p->X, ypx);
    mul2(Cc, D, p->T, q2dT, p->Z, q->Z);
    E = B - A; Fv = D - Cc; G = D + Cc; H = B + A;
  } else {
    F qX = cneg(q->X, negate), qdT = cneg(q->dT, negate);
    mul2(A, B, p->X, qX, p->Y, q->Y);
    mul2(Cc, D, p->T, qdT, p->Z, q->Z);
    E = (p->X + p->Y) * (qX + q->Y) - A - B;
    Fv = D - Cc; G = D + Cc; H = B + C::mul_neg_a(A);
  }
  mul2(X3, Y3, E, Fv, G, H);
  r->X = X3; r->Y = Y3;
  if (want_t) { F T3; mul2(T3, Z3, E, H, Fv, G); r->T = T3; r->Z = Z3; } else r->Z = Fv * G;
}
// r = p + q, q affine cached, optionally negated (8M)
template <class C> HD_NOINLINE void te_madd(TEPoint<C>* r, const TEPoint<C>* p, const TEAffCached<C>* q, bool negate, bool want_t = true) {
  typedef typename C::F F;
  F A, B, Cc, E, D, Fv, G, H, X3, Y3, Z3;
  if constexpr (C::A_IS_M1) {
    F ypx = select(negate, q->y, q->x), ymx = select(negate, q->x, q->y), q2dt = cneg(q->dt, negate);
    mul2(A, B, p->Y - p->X, ymx, p->Y + p->X, ypx);
    Cc = p->T * q2dt;
    D = dbl(p->Z);
    E = B - A; Fv = D - Cc; G = D + Cc; H = B + A;
  } else {
    F qx = cneg(q->x, negate), qdt = cneg(q->dt, negate);
    D = p->Z;
    mul2(A, B, p->X, qx, p->Y, q->y);
    mul2(Cc, E, p->T, qdt, p->X + p->Y, qx + q->y);
    E = E - A - B;
    Fv = D - Cc; G = D + Cc; H = B + C::mul_neg_a(A);
  }
  mul2(X3, Y3, E, Fv, G, H);
  r->X = X3; r->Y = Y3;
  if (want_t) { F T3; mul2(T3, Z3, E, H, Fv, G); r->T = T3; r->Z = Z3; } else r->Z = Fv * G;
}
// r = 2p (4S + 4M; T of the result is skipped when the next operation is another doubling).
// dbl-2008-hwcd with D = a A, G = D + B, F = G - C, H = D - B.  Computed with Dn = -a A (a = -5: 5A, three additions and no negation),
// G = B - Dn and the NEGATED F and H (Fv = C - G, H = Dn + B): all four coordinates come out negated, which is the same point.
template <class C> HD_NOINLINE void te_dbl(TEPoint<C>* r, const TEPoint<C>* p, bool want_t) {
  typedef typename C::F F;
  F A, B, Cc, E, X3, Y3, Z3;
  sqr2(A, B, p->X, p->Y);
  sqr2(Cc, E, p->Z, p->X + p->Y);
  Cc = dbl(Cc);
  F Dn = C::mul_neg_a(A);
  E = E - A - B;
  F G = B - Dn, Fv = Cc - G, H = Dn + B;
  mul2(X3, Y3, E, Fv, G, H);
  r->X = X3; r->Y = Y3;
  if (want_t) { F T3; mul2(Z3, T3, Fv, G, E, H); r->Z = Z3; r->T = T3; } else r->Z = Fv * G;
}
template <class C> HD_INLINE void te_neg(TEPoint<C>& P) { P.X = neg(P.X); P.T = neg(P.T); }

// ---- Bandersnatch GLV endomorphism psi(P) = lambda*P on the prime-order subgroup:
// psi(x,y) = (c(1-y^2)/(xy), b(y^2+b)/(y^2-b));  projectively (f*h : g*XY : XY*h : f*g) with
// f = c(Z^2-Y^2), g = b(Y^2+bZ^2), h = Y^2-bZ^2.  x = 0 (identity / the 2-torsion point) maps to identity.
HD_NOINLINE void band_endo(TEPoint<BandCurve>* r, const TEPoint<BandCurve>* p) {
  typedef BandCurve::F F;
  F b = fconst<BlsFr, BandConsts::ENDO_B>(), c = fconst<BlsFr, BandConsts::ENDO_C>();
  F yy = sqr(p->Y), zz = sqr(p->Z), xy = p->X * p->Y, bzz = b * zz;
  F f = c * (zz - yy), g = b * (yy + bzz), h = yy - bzz;
  bool degenerate = p->X.is_zero();
  TEPoint<BandCurve> id; te_set_identity(id);
  r->X = select(degenerate, id.X, f * h);
  r->Y = select(degenerate, id.Y, g * xy);
  r->Z = select(degenerate, id.Z, xy * h);
  r->T = select(degenerate, id.T, f * g);
}

}  // namespace vrfs
