// K6 of SURVEY.md 2.5: Suite::data_to_point - Elligator2 (Bandersnatch: `utils::hash_to_curve_ell2_rfc_9380`
// over ark-ec's Elligator2Map + ark-ff's DefaultFieldHasher) and try-and-increment
// (`utils::hash_to_curve_tai_rfc_9381`), both named at /root/reference/src/lib.rs:13-17; exact byte
// layouts in SURVEY.md A.5 (incl. ark-ff's 48-byte Z_pad).  Also codec point_decode (A.2).
#pragma once
#include "suite.cuh"

namespace vrfs {

struct ExpTM1H { template <class P> static HD_INLINE uint32_t get(int i) { return P::tm1h(i); } };

// Square root with warp-uniform control flow (Tonelli-Shanks with the 2-adic part resolved by
// conditional moves).  Returns false (out unspecified) if a is not a square.  WHICH root is returned is
// irrelevant to callers: every use fixes the sign afterwards (parity / "is_high" flag).
template <class P>
HD_NOINLINE bool sqrt_ct(Fp<P>* out, const Fp<P>* a_) {
  typedef Fp<P> F;
  const F a = *a_;
  F w = pow_const<P, ExpTM1H>(a);       // a^((t-1)/2)
  F z = w * a;                          // a^((t+1)/2)
  F tt = z * w;                         // a^t, lies in the 2^s-torsion
  F c = fconst<P, P::rou>();
  F b = tt;
  for (int i = P::TWO_ADICITY; i >= 2; i--) {
    for (int j = 1; j <= i - 2; j++) b = sqr(b);
    bool e = (b == F::one());
    F zt = z * c;
    z = select(e, z, zt);
    c = sqr(c);
    F t2 = tt * c;
    tt = select(e, tt, t2);
    b = tt;
  }
  *out = z;
  return sqr(z) == a;
}

// ---- Elligator2 for Bandersnatch (A.5): Montgomery J = A, K = B, Z = 5; result in extended TE coordinates
HD_NOINLINE void band_elligator2(TEPoint<BandCurve>* out, const Fp<BlsFr>* u_) {
  typedef Fp<BlsFr> F;
  typedef BandConsts K;
  const F one = F::one(), JK = fconst<BlsFr, K::ELL2_JK>(), KSQI = fconst<BlsFr, K::ELL2_KSQI>(), Kc = fconst<BlsFr, K::ELL2_K>();
  F u = *u_;
  F uu = sqr(u);
  F den = dbl(dbl(uu)) + uu + one;                   // 1 + Z*u^2, Z = 5
  if (den.is_zero()) den = one;
  F x1 = neg(JK * inv(den));
  F t = sqr(x1);
  F gx1 = t * x1 + t * JK + x1 * KSQI;               // g(x) = x^3 + (J/K) x^2 + x/K^2
  F x2 = neg(x1) - JK;
  t = sqr(x2);
  F gx2 = t * x2 + t * JK + x2 * KSQI;
  F y1, y2;
  bool sq1 = gx1.is_zero() | sqrt_ct<BlsFr>(&y1, &gx1);
  if (gx1.is_zero()) y1 = F::zero();
  bool sq2 = sqrt_ct<BlsFr>(&y2, &gx2);
  (void)sq2;                                          // exactly one of gx1, gx2 is a square (Z non-square)
  F x = select(sq1, x1, x2), y = select(sq1, y1, y2);
  if (is_odd(y) != sq1) y = neg(y);                   // sgn0(y) = 1 on the first branch, 0 on the second
  F s = x * Kc, tm = y * Kc;                          // Montgomery (s, t) -> TE (s/t, (s-1)/(s+1))
  F sp1 = s + one, sm1 = s - one;
  F Z = tm * sp1;
  bool degenerate = Z.is_zero();
  TEPoint<BandCurve> id; te_set_identity(id);
  out->X = select(degenerate, id.X, s * sp1);
  out->Y = select(degenerate, id.Y, sm1 * tm);
  out->Z = select(degenerate, id.Z, Z);
  out->T = select(degenerate, id.T, s * sm1);
}

// hash_to_curve_ell2_rfc_9380: DST = "ECVRF_" || h2c_id || SUITE_ID, expand_message_xmd (SHA-512, 96 bytes,
// ark-ff's 48-byte Z_pad), two field elements, two Elligator2 maps, add, clear cofactor.
HD_INLINE void band_h2c_ell2(TEPoint<BandCurve>& P, const uint8_t* data, uint32_t len) {
  constexpr char dst[] = "ECVRF_Bandersnatch_XMD:SHA-512_ELL2_RO_Bandersnatch_SHA-512_ELL2";
  constexpr int DL = sizeof(dst) - 1;   // 64
  uint8_t b0[64], b1[64], b2[64];
  Sha512 h; h.init();
  for (int i = 0; i < 48; i++) h.put(0);
  h.update(data, len);
  h.put(0x00); h.put(0x60); h.put(0x00);
  for (int i = 0; i < DL; i++) h.put((uint8_t)dst[i]);
  h.put((uint8_t)DL); h.final(b0);
  h.init(); h.update(b0, 64); h.put(0x01);
  for (int i = 0; i < DL; i++) h.put((uint8_t)dst[i]);
  h.put((uint8_t)DL); h.final(b1);
  h.init();
  for (int i = 0; i < 64; i++) h.put(b0[i] ^ b1[i]);
  h.put(0x02);
  for (int i = 0; i < DL; i++) h.put((uint8_t)dst[i]);
  h.put((uint8_t)DL); h.final(b2);
  Fp<BlsFr> u[2];
  for (int j = 0; j < 2; j++) {          // u_j = BE(uniform[48j .. 48j+48]) mod q, uniform = b1 || b2
    uint32_t lo[8], hi[8];
    for (int i = 0; i < 8; i++) { lo[i] = 0; hi[i] = 0; }
    for (int i = 0; i < 48; i++) {
      int src = 48 * j + i;
      uint32_t byte = src < 64 ? b1[src] : b2[src - 64];
      int pos = 47 - i;
      if (pos < 32) lo[pos >> 2] |= byte << (8 * (pos & 3)); else hi[(pos - 32) >> 2] |= byte << (8 * (pos & 3));
    }
    u[j] = to_mont_wide<BlsFr>(lo, hi);
  }
  TEPoint<BandCurve> Q;
  band_elligator2(&P, &u[0]);
  band_elligator2(&Q, &u[1]);
  te_add<BandCurve>(&P, &P, &Q);
  te_dbl<BandCurve>(&P, &P, true); te_dbl<BandCurve>(&P, &P, true);
}

// ArkworksCodec point_decode for TE curves (A.2): no subgroup check.  Montgomery-form affine out.
template <class C>
HD_INLINE bool ark_decode_point(typename C::F& x, typename C::F& y, const uint8_t* in) {
  typedef typename C::F F;
  uint32_t raw[8];
  load_le<8>(raw, in);
  bool sign = raw[7] >> 31;
  raw[7] &= 0x7fffffffu;
  if (!is_canonical<typename C::Fq>(raw)) return false;
  y = to_mont<typename C::Fq>(raw);
  F yy = sqr(y), num = F::one() - yy, den = C::mul_a(F::one()) - C::d() * yy;
  if (den.is_zero()) return false;
  F x2 = num * inv(den);
  if (x2.is_zero()) x = F::zero();
  else if (!sqrt_ct<typename C::Fq>(&x, &x2)) return false;
  if (is_high(x) != sign) x = neg(x);
  return true;
}

// Sec1Codec point_decode (A.2): (0x02|0x03) || BE32(x); y = sqrt(x^3 + a x + b) with the requested parity
template <class C>
HD_INLINE bool sec1_decode_point(typename C::F& x, typename C::F& y, const uint8_t* in) {
  typedef typename C::F F;
  if (in[0] != 2 && in[0] != 3) return false;
  uint32_t raw[8];
  load_be<8>(raw, in + 1);
  if (!is_canonical<typename C::Fq>(raw)) return false;
  x = to_mont<typename C::Fq>(raw);
  F rhs = (sqr(x) + C::mul_a(F::one())) * x + C::b();
  if (rhs.is_zero()) y = F::zero();
  else if (!sqrt_ct<typename C::Fq>(&y, &rhs)) return false;
  if (is_odd(y) != (bool)(in[0] & 1)) y = neg(y);
  return true;
}
template <class S> HD_INLINE bool decode_point(typename S::C::F& x, typename S::C::F& y, const uint8_t* in) {
  if constexpr (S::SEC1) return sec1_decode_point<typename S::C>(x, y, in); else return ark_decode_point<typename S::C>(x, y, in);
}

// hash_to_curve_tai_rfc_9381 (A.5): first ctr whose hash decodes to a point whose cofactor multiple is not
// the identity.  LE codec: decode hs[0..32]; SEC1 codec: decode 0x02 || hs.
template <class S>
HD_INLINE bool h2c_tai(typename Grp<typename S::C>::Pt& P, const uint8_t* data, uint32_t len) {
  typedef typename S::C C;
  typedef Grp<C> G;
  for (int ctr = 0; ctr < 256; ctr++) {
    typename S::H h; h.init();
    put_suite_id<S>(h); h.put(0x01); h.update(data, len); h.put((uint8_t)ctr); h.put(0x00);
    uint8_t hs[S::HLEN + 1];
    hs[0] = 0x02;
    h.final(hs + 1);
    typename C::F x, y;
    if (!decode_point<S>(x, y, S::SEC1 ? hs : hs + 1)) continue;
    G::from_affine(P, x, y);
    for (int i = 0; i < C::COF_LOG2; i++) G::dbl(&P);
    if constexpr (C::IS_TE) { if (te_is_identity<C>(P)) continue; }
    return true;
  }
  return false;
}

}  // namespace vrfs
