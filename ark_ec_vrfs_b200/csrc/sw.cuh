// K4 of SURVEY.md 2.5: short-Weierstrass points y^2 = x^3 + a*x + b in homogeneous projective
// coordinates (X:Y:Z), x = X/Z, y = Y/Z, identity (0:1:0) - the device replacement for ark-ec's
// `short_weierstrass::Projective` behind `AffinePoint<S>` (/root/reference/src/lib.rs:13-17) for secp256r1
// (a = -3) and BLS12-381 G1 (a = 0).
// Addition is the COMPLETE formula of Renes-Costello-Batina (EUROCRYPT 2016, Algorithm 1): one
// branch-free sequence valid for every pair of inputs (P+Q, P+P, P+(-P), identity), so a warp never
// diverges on exceptional cases and adversarial verify inputs need no special handling.  (The reference
// uses Jacobian formulas with explicit case analysis; the group element computed is the same.)
#pragma once
#include "te.cuh"

namespace vrfs {

struct P256Curve {
  typedef P256Fp Fq;
  typedef P256Fr Fr;
  typedef P256Consts K;
  typedef Fp<Fq> F;
  static constexpr bool IS_TE = false;
  static constexpr bool HAS_GLV = false;
  static constexpr bool A_IS_M3 = true;
  static constexpr int COF_LOG2 = 0;
  static HD_INLINE F mul_a(const F& x) { return neg(dbl(x) + x); }                        // a = -3
  static HD_INLINE F mul_b3(const F& x) { return x * fconst<Fq, K::B3>(); }
  static HD_INLINE F b() { return fconst<Fq, K::B>(); }
  static HD_INLINE F gx() { return fconst<Fq, K::GX>(); }
  static HD_INLINE F gy() { return fconst<Fq, K::GY>(); }
  static HD_INLINE F bx() { return fconst<Fq, K::BX>(); }
  static HD_INLINE F by() { return fconst<Fq, K::BY>(); }
};
// Bandersnatch in short-Weierstrass form (ark-ed-on-bls12-381-bandersnatch SWConfig): general a, cofactor 4
struct BandSwCurve {
  typedef BlsFr Fq;
  typedef BandFr Fr;
  typedef BandSwConsts K;
  typedef Fp<Fq> F;
  static constexpr bool IS_TE = false;
  static constexpr bool HAS_GLV = false;
  static constexpr bool A_IS_M3 = false;
  static constexpr int COF_LOG2 = 2;
  static HD_INLINE F mul_a(const F& x) { return x * fconst<Fq, K::A>(); }
  static HD_INLINE F mul_b3(const F& x) { return x * fconst<Fq, K::B3>(); }
  static HD_INLINE F b() { return fconst<Fq, K::B>(); }
  static HD_INLINE F gx() { return fconst<Fq, K::GX>(); }
  static HD_INLINE F gy() { return fconst<Fq, K::GY>(); }
  static HD_INLINE F bx() { return fconst<Fq, K::BX>(); }
  static HD_INLINE F by() { return fconst<Fq, K::BY>(); }
};
struct G1Curve {
  typedef BlsFq Fq;
  typedef BlsFr Fr;
  typedef G1Consts K;
  typedef Fp<Fq> F;
  static constexpr bool IS_TE = false;
  static constexpr bool HAS_GLV = false;
  static constexpr int COF_LOG2 = 0;
  static HD_INLINE F mul_a(const F&) { return F::zero(); }                                 // a = 0
  static HD_INLINE F mul_b3(const F& x) { F x4 = dbl(dbl(x)); return dbl(x4) + x4; }       // 3b = 12
  static HD_INLINE F b() { return fconst<Fq, K::B>(); }
  static HD_INLINE F gx() { return fconst<Fq, K::GX>(); }
  static HD_INLINE F gy() { return fconst<Fq, K::GY>(); }
};

template <class C> struct SWPoint { typename C::F X, Y, Z; };

template <class C> HD_INLINE void sw_set_identity(SWPoint<C>& P) { P.X = C::F::zero(); P.Y = C::F::one(); P.Z = C::F::zero(); }
template <class C> HD_INLINE void sw_from_affine(SWPoint<C>& P, const typename C::F& x, const typename C::F& y) { P.X = x; P.Y = y; P.Z = C::F::one(); }
template <class C> HD_INLINE bool sw_is_identity(const SWPoint<C>& P) { return P.Z.is_zero(); }
template <class C> HD_INLINE bool sw_on_curve(const typename C::F& x, const typename C::F& y) {
  return sqr(y) == (sqr(x) + C::mul_a(C::F::one())) * x + C::b();
}
template <class C> HD_INLINE void sw_cneg(SWPoint<C>& P, bool c) { P.Y = cneg(P.Y, c); }

// r = p + q, complete (RCB16 Algorithm 1; 12M + 3 mul_a + 2 mul_b3)
template <class C> HD_NOINLINE void sw_add(SWPoint<C>* r, const SWPoint<C>* p, const SWPoint<C>* q) {
  typedef typename C::F F;
  F t0 = p->X * q->X, t1 = p->Y * q->Y, t2 = p->Z * q->Z;
  F t3 = (p->X + p->Y) * (q->X + q->Y) - (t0 + t1);          // X1Y2 + X2Y1
  F t4 = (p->X + p->Z) * (q->X + q->Z) - (t0 + t2);          // X1Z2 + X2Z1
  F t5 = (p->Y + p->Z) * (q->Y + q->Z) - (t1 + t2);          // Y1Z2 + Y2Z1
  F Z3 = C::mul_b3(t2) + C::mul_a(t4);
  F X3 = t1 - Z3;
  Z3 = t1 + Z3;
  F Y3 = X3 * Z3;
  t1 = dbl(t0) + t0;
  F at2 = C::mul_a(t2);
  t4 = C::mul_b3(t4);
  t1 = t1 + at2;
  t2 = C::mul_a(t0 - at2);
  t4 = t4 + t2;
  Y3 = Y3 + t1 * t4;
  X3 = t3 * X3 - t5 * t4;
  Z3 = t5 * Z3 + t3 * t1;
  r->X = X3; r->Y = Y3; r->Z = Z3;
}

// r = 16 p for a = -3: homogeneous -> Jacobian (X Z : Y Z^2 : Z), four Jacobian doublings (dbl-2001-b, 3M + 5S each, no
// exceptional case on a prime-order curve), back to homogeneous (X Z : Y : Z^3): 38 products instead of the 56 of four complete
// additions.  The identity (0:1:0) passes through as (0:0:0) and is restored at the end.
template <class C> HD_NOINLINE void sw_dbl4_am3(SWPoint<C>* r, const SWPoint<C>* p) {
  typedef typename C::F F;
  F Z = p->Z, zz = sqr(Z), X = p->X * Z, Y = p->Y * zz;
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
    F delta = sqr(Z), gamma = sqr(Y), beta = X * gamma;
    F t = (X - delta) * (X + delta), alpha = dbl(t) + t;
    F b4 = dbl(dbl(beta));
    F X3 = sqr(alpha) - dbl(b4);
    F Z3 = sqr(Y + Z) - gamma - delta;
    F g2 = sqr(gamma), g8 = dbl(dbl(dbl(g2)));
    Y = alpha * (b4 - X3) - g8;
    X = X3; Z = Z3;
  }
  bool inf = Z.is_zero();
  F z2 = sqr(Z);
  r->X = X * Z;
  r->Y = select(inf, F::one(), Y);
  r->Z = z2 * Z;
}

}  // namespace vrfs
