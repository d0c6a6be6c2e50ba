// K6 of SURVEY.md 2.5: Suite::data_to_point - Elligator2 (Bandersnatch: `utils::hash_to_curve_ell2_rfc_9380`
// over ark-ec's Elligator2Map + ark-ff's DefaultFieldHasher) and try-and-increment
// (`utils::hash_to_curve_tai_rfc_9381`), both named at /root/reference/src/lib.rs:13-17; exact byte
// layouts in SURVEY.md A.5 (incl. ark-ff's 48-byte Z_pad).  Also codec point_decode (A.2).
#pragma once
#include <type_traits>
#include "suite.cuh"
#include "gen/bls_torsion.cuh"

namespace vrfs {

struct ExpTM1H { template <class P> static HD_INLINE uint32_t get(int i) { return P::tm1h(i); } };

// Square root with warp-uniform control flow (Tonelli-Shanks with the 2-adic part resolved by
// conditional moves).  Returns false (out unspecified) if a is not a square.  WHICH root is returned is
// irrelevant to callers: every use fixes the sign afterwards (parity / "is_high" flag).
template <class P>
HD_NOINLINE bool sqrt_ts(Fp<P>* out, const Fp<P>* a_) {
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

// ---- BLS12-381 Fr (Bandersnatch base field), q - 1 = 2^32 t: the 2^32-torsion part of a^t is resolved by a
// Pohlig-Hellman discrete logarithm over four 8-bit windows with precomputed tables (gen/bls_torsion.cuh) instead of
// Tonelli-Shanks' quadratic loop: with g = 5^t, a^t = g^k,  a square <=> k even,  sqrt(a) = a^((t+1)/2) g^(-k/2),
// 1/a = (a^((t-1)/2))^2 g^(-k).  One exponentiation therefore yields the square root, the quadratic character AND
// the inverse - which makes Elligator2 and point decompression inversion-free.
HD_INLINE Fp<BlsFr> bls_torsion_w(int i, uint32_t j) {
  Fp<BlsFr> r;
  const uint4* p = reinterpret_cast<const uint4*>(&BLS_TORSION_W[i][j][0]);
  uint4 a = p[0], b = p[1];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
HD_INLINE uint32_t bls_torsion_lookup(const Fp<BlsFr>& x) {    // x = g^(2^24 j) -> j
  return BLS_TORSION_HASH[(x.v[0] * BLS_TORSION_HASH_MUL) >> BLS_TORSION_HASH_SHIFT];
}
// k with tt = g^k, tt in the 2^32-torsion
HD_NOINLINE uint32_t bls_torsion_dlog(const Fp<BlsFr>* tt) {
  typedef Fp<BlsFr> F;
  F y0 = *tt, y1 = y0;
#pragma unroll 1
  for (int i = 0; i < 8; i++) y1 = sqr(y1);
  F y2 = y1;
#pragma unroll 1
  for (int i = 0; i < 8; i++) y2 = sqr(y2);
  F y3 = y2;
#pragma unroll 1
  for (int i = 0; i < 8; i++) y3 = sqr(y3);
  uint32_t k0 = bls_torsion_lookup(y3);
  uint32_t k1 = bls_torsion_lookup(y2 * bls_torsion_w(2, k0));
  uint32_t k2 = bls_torsion_lookup(y1 * bls_torsion_w(1, k0) * bls_torsion_w(2, k1));
  uint32_t k3 = bls_torsion_lookup(y0 * bls_torsion_w(0, k0) * bls_torsion_w(1, k1) * bls_torsion_w(2, k2));
  return k0 | (k1 << 8) | (k2 << 16) | (k3 << 24);
}
// g^(-m)
HD_NOINLINE Fp<BlsFr> bls_torsion_pow_neg(uint32_t m) {
  return bls_torsion_w(0, m & 255u) * bls_torsion_w(1, (m >> 8) & 255u) * bls_torsion_w(2, (m >> 16) & 255u) * bls_torsion_w(3, m >> 24);
}
// a != 0:  w = a^((t-1)/2),  z = a^((t+1)/2),  k = log_g(a^t)
HD_INLINE void bls_sqrt_parts(Fp<BlsFr>& w, Fp<BlsFr>& z, uint32_t& k, const Fp<BlsFr>& a) {
  w = pow_const<BlsFr, ExpTM1H>(a);
  z = w * a;
  Fp<BlsFr> tt = z * w;
  k = bls_torsion_dlog(&tt);
}
HD_NOINLINE bool bls_sqrt(Fp<BlsFr>* out, const Fp<BlsFr>* a_) {
  typedef Fp<BlsFr> F;
  const F a = *a_;
  if (a.is_zero()) { *out = F::zero(); return true; }
  F w, z; uint32_t k;
  bls_sqrt_parts(w, z, k, a);
  if (k & 1u) return false;
  *out = z * bls_torsion_pow_neg(k >> 1);
  return sqr(*out) == a;
}
template <class P> HD_INLINE bool sqrt_ct(Fp<P>* out, const Fp<P>* a) {
  if constexpr (std::is_same<P, BlsFr>::value) return bls_sqrt(out, a); else return sqrt_ts<P>(out, a);
}
// x with x^2 = num/den (den != 0), no inversion: false if num/den is not a square.  Also usable for other fields (generic path).
template <class P> HD_INLINE bool sqrt_ratio(Fp<P>* x, const Fp<P>& num, const Fp<P>& den) {
  Fp<P> x2 = num * inv(den);
  if (x2.is_zero()) { *x = Fp<P>::zero(); return true; }
  return sqrt_ct<P>(x, &x2);
}
template <> HD_INLINE bool sqrt_ratio<BlsFr>(Fp<BlsFr>* x, const Fp<BlsFr>& num, const Fp<BlsFr>& den) {
  typedef Fp<BlsFr> F;
  if (num.is_zero()) { *x = F::zero(); return true; }
  F a = num * den, w, z; uint32_t k;
  bls_sqrt_parts(w, z, k, a);
  if (k & 1u) return false;
  F r = z * bls_torsion_pow_neg(k >> 1);              // sqrt(num den)
  F ia = sqr(w) * bls_torsion_pow_neg(k);             // 1/(num den)
  *x = r * num * ia;                                  // sqrt(num den)/den
  return sqr(*x) * den == num;
}

// ---- Elligator2 for Bandersnatch (A.5): Montgomery J = A, K = B, Z = 5; result in extended TE coordinates
HD_NOINLINE void band_elligator2(TEPoint<BandCurve>* out, const Fp<BlsFr>* u_) {
  typedef Fp<BlsFr> F;
  typedef BandConsts K;
  const F one = F::one(), JK = fconst<BlsFr, K::ELL2_JK>(), KSQI = fconst<BlsFr, K::ELL2_KSQI>(), Kc = fconst<BlsFr, K::ELL2_K>();
  F u = *u_;
  F uu = sqr(u);
  F D = dbl(dbl(uu)) + uu + one;                     // 1 + Z*u^2, Z = 5 (never 0: -1/5 is not a square)
  if (D.is_zero()) D = one;
  // x1 = N/D with N = -J/K;  g(x1) = G1/D^4 with G1 = N (N^2 + (J/K) N D + D^2/K^2) D, so sqrt(g(x1)) = sqrt(G1)/D^2 and
  // 1/D = N (N^2 + ...)/G1: one exponentiation (bls_sqrt_parts) gives the root, the character and the inverse.
  const F N = neg(JK);
  F ND = N * D, NP = N * (sqr(N) + JK * ND + KSQI * sqr(D));
  F G1 = NP * D;
  F x, y;
  bool sq1;
  if (G1.is_zero()) {                                // x1 is a root of g: y = 0 (measure-zero case, plain inversion)
    sq1 = true; x = N * inv(D); y = F::zero();
  } else {
    F w, z; uint32_t k;
    bls_sqrt_parts(w, z, k, G1);
    sq1 = !(k & 1u);
    F invD = NP * sqr(w) * bls_torsion_pow_neg(k);   // NP / G1
    F x1 = N * invD;
    // k even: sqrt(G1) = z g^(-k/2).  k odd: g(x2) = Z u^2 g(x1) and (Z G1)^t = g^(k+1) (Z = 5 generates the torsion part),
    // so sqrt(g(x2)) = u Z^((t+1)/2) z g^(-(k+1)/2) / D^2
    uint32_t m = (uint32_t)(((unsigned long long)k + (k & 1u)) >> 1);
    F r = z * bls_torsion_pow_neg(m);
    if (!sq1) r = r * fconst<BlsFr, K::ELL2_C2>() * u;
    y = r * sqr(invD);
    x = sq1 ? x1 : neg(x1) - JK;
  }
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
  if (!sqrt_ratio<typename C::Fq>(&x, num, den)) return false;
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
// arkworks codec on a short-Weierstrass curve [RECALL]: 32-byte LE x, then a flag byte whose bit 7 selects the larger root and whose
// bit 6 marks the identity (rejected here: no typed Public / Input / Output holds it)
template <class C>
HD_INLINE bool arksw_decode_point(typename C::F& x, typename C::F& y, const uint8_t* in) {
  typedef typename C::F F;
  const uint32_t flags = in[32] >> 6;
  if (flags & 1u) return false;
  uint32_t raw[8];
  load_le<8>(raw, in);
  if (!is_canonical<typename C::Fq>(raw)) return false;
  x = to_mont<typename C::Fq>(raw);
  F rhs = (sqr(x) + C::mul_a(F::one())) * x + C::b();
  if (rhs.is_zero()) y = F::zero();
  else if (!sqrt_ct<typename C::Fq>(&y, &rhs)) return false;
  if (is_high(y) != (bool)(flags >> 1)) y = neg(y);
  return true;
}
template <class S> HD_INLINE bool decode_point(typename S::C::F& x, typename S::C::F& y, const uint8_t* in) {
  if constexpr (S::SEC1) return sec1_decode_point<typename S::C>(x, y, in);
  else if constexpr (S::ARK_SW) return arksw_decode_point<typename S::C>(x, y, in);
  else return ark_decode_point<typename S::C>(x, y, in);
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
