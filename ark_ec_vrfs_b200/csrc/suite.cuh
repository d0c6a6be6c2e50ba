// Suite procedures (K5/K6/K9-K11 glue of SURVEY.md 2.5): the device counterpart of the `Suite` trait and
// `utils::{challenge_rfc_9381, nonce_rfc_8032, point_to_hash_rfc_9381}` + `codec::ArkworksCodec`, all
// named at /root/reference/src/lib.rs:13-17 and specified in SURVEY.md Appendix A (A.2, A.6-A.10).
#pragma once
#include "te_lincomb.cuh"
#include "sha2.cuh"

namespace vrfs {

struct BandSuite {
  typedef BandCurve C;
  typedef Sha512 H;
  static constexpr int CLEN = 32, HLEN = 64, ID_LEN = 25;
  static HD_INLINE uint8_t id(int i) { constexpr char s[] = "Bandersnatch_SHA-512_ELL2"; return (uint8_t)s[i]; }
};
struct EdSuite {
  typedef EdCurve C;
  typedef Sha512 H;
  static constexpr int CLEN = 16, HLEN = 64, ID_LEN = 19;
  static HD_INLINE uint8_t id(int i) { constexpr char s[] = "Ed25519_SHA-512_TAI"; return (uint8_t)s[i]; }
};

template <class S> HD_INLINE void put_suite_id(typename S::H& h) { for (int i = 0; i < S::ID_LEN; i++) h.put(S::id(i)); }

// ArkworksCodec point_encode (A.2): 32-byte LE y, bit 255 set iff x > (p-1)/2.  x, y canonical limbs.
template <class C> HD_INLINE void ark_encode_point(uint8_t* out, const uint32_t* x, const uint32_t* y) {
  uint32_t h[8], t[8];
  for (int i = 0; i < 8; i++) h[i] = C::Fq::pm1h(i);
  bool high = MontChains<8>::sub(t, h, x) != 0;
  store_le<8>(out, y);
  if (high) out[31] |= 0x80;
}
template <class C> HD_INLINE void ark_encode_point_mont(uint8_t* out, const typename C::F& x, const typename C::F& y) {
  uint32_t rx[8], ry[8];
  from_mont<typename C::Fq>(rx, x); from_mont<typename C::Fq>(ry, y);
  ark_encode_point<C>(out, rx, ry);
}
// encode straight from the ABI's affine bytes (x||y LE canonical, already validated)
template <class C> HD_INLINE void ark_encode_point_bytes(uint8_t* out, const uint8_t* p) {
  uint32_t rx[8], ry[8];
  load_le<8>(rx, p); load_le<8>(ry, p + 32);
  ark_encode_point<C>(out, rx, ry);
}

// hash output -> scalar mod r, canonical limbs.  nbytes = 16, 32 or 64; big- or little-endian (A.6, A.7, A.10)
template <class C> HD_INLINE void hash_to_scalar(uint32_t* k, const uint8_t* d, int nbytes, bool big_endian) {
  uint32_t lo[8], hi[8];
  for (int i = 0; i < 8; i++) { lo[i] = 0; hi[i] = 0; }
  for (int i = 0; i < nbytes; i++) {
    int pos = big_endian ? nbytes - 1 - i : i;       // byte i has weight 256^pos
    uint32_t b = d[i];
    if (pos < 32) lo[pos >> 2] |= b << (8 * (pos & 3)); else hi[(pos - 32) >> 2] |= b << (8 * (pos & 3));
  }
  Fp<typename C::Fr> m = nbytes > 32 ? to_mont_wide<typename C::Fr>(lo, hi) : to_mont<typename C::Fr>(lo);
  from_mont<typename C::Fr>(k, m);
}

// challenge_rfc_9381 (A.7): c = BE(H(suite || 0x02 || enc(P1..P5) || ad || 0x00)[..cLen]) mod r
template <class S> HD_INLINE void suite_challenge(uint32_t* c, const uint8_t (*enc)[32], const uint8_t* ad, uint32_t adlen) {
  typename S::H h; h.init();
  put_suite_id<S>(h); h.put(0x02);
  for (int p = 0; p < 5; p++) h.update(enc[p], 32);
  h.update(ad, adlen); h.put(0x00);
  uint8_t dig[S::HLEN]; h.final(dig);
  hash_to_scalar<typename S::C>(c, dig, S::CLEN, true);
}
// point_to_hash_rfc_9381 (A.8)
template <class S> HD_INLINE void suite_point_to_hash(uint8_t* out, const uint8_t* enc) {
  typename S::H h; h.init();
  put_suite_id<S>(h); h.put(0x03); h.update(enc, 32); h.put(0x00); h.final(out);
}
// nonce_rfc_8032 (A.6): k = LE(H(H(enc_sc(sk))[32..64] || enc_pt(I))) mod r
template <class S> HD_INLINE void suite_nonce_8032(uint32_t* k, const uint32_t* sk, const uint8_t* enc_input) {
  uint8_t e[32], d[64];
  store_le<8>(e, sk);
  typename S::H h; h.init(); h.update(e, 32); h.final(d);
  h.init(); h.update(d + 32, 32); h.update(enc_input, 32); h.final(d);
  hash_to_scalar<typename S::C>(k, d, 64, false);
}
// Pedersen blinding (A.10): b = BE(H(suite || 0xCC || enc_sc(sk) || enc_pt(I) || ad || 0x00)) mod r
template <class S> HD_INLINE void suite_blinding(uint32_t* b, const uint32_t* sk, const uint8_t* enc_input, const uint8_t* ad, uint32_t adlen) {
  uint8_t e[32], d[64];
  store_le<8>(e, sk);
  typename S::H h; h.init();
  put_suite_id<S>(h); h.put(0xCC); h.update(e, 32); h.update(enc_input, 32); h.update(ad, adlen); h.put(0x00); h.final(d);
  hash_to_scalar<typename S::C>(b, d, S::HLEN, true);
}
// s = k + c*x mod r (canonical limbs in and out)
template <class C> HD_INLINE void scalar_muladd(uint32_t* s, const uint32_t* k, const uint32_t* c, const uint32_t* x) {
  typedef typename C::Fr Fr;
  Fp<Fr> km = to_mont<Fr>(k), cm = to_mont<Fr>(c), xm = to_mont<Fr>(x);
  from_mont<Fr>(s, km + cm * xm);
}

// two projective points (X,Y,Z Montgomery limbs) -> affine with one shared inversion
template <class C> HD_INLINE void te_two_to_affine(typename C::F* ax, typename C::F* ay, const uint32_t* p0, const uint32_t* p1) {
  typedef typename C::F F;
  F X0, Y0, Z0, X1, Y1, Z1;
  for (int i = 0; i < 8; i++) { X0.v[i] = p0[i]; Y0.v[i] = p0[8 + i]; Z0.v[i] = p0[16 + i]; X1.v[i] = p1[i]; Y1.v[i] = p1[8 + i]; Z1.v[i] = p1[16 + i]; }
  F ti = inv(Z0 * Z1);
  F z0i = ti * Z1, z1i = ti * Z0;
  ax[0] = X0 * z0i; ay[0] = Y0 * z0i; ax[1] = X1 * z1i; ay[1] = Y1 * z1i;
}

// ietf::Verifier::verify, second half (A.9): U, V already computed; accept iff challenge(Y,I,O,U,V,ad) == c
template <class S>
HD_INLINE bool ietf_verify_finish_item(const uint8_t* pk, const uint8_t* input, const uint8_t* output, const uint8_t* c_bytes,
                                       const uint32_t* u_xyz, const uint32_t* v_xyz, const uint8_t* ad, uint32_t adlen) {
  typedef typename S::C C;
  typename C::F ax[2], ay[2];
  te_two_to_affine<C>(ax, ay, u_xyz, v_xyz);
  uint8_t enc[5][32];
  ark_encode_point_bytes<C>(enc[0], pk);
  ark_encode_point_bytes<C>(enc[1], input);
  ark_encode_point_bytes<C>(enc[2], output);
  ark_encode_point_mont<C>(enc[3], ax[0], ay[0]);
  ark_encode_point_mont<C>(enc[4], ax[1], ay[1]);
  uint32_t c2[8], c[8];
  suite_challenge<S>(c2, enc, ad, adlen);
  hash_to_scalar<C>(c, c_bytes, 32, false);
  uint32_t diff = 0;
  for (int i = 0; i < 8; i++) diff |= c[i] ^ c2[i];
  return diff == 0;
}

// K projective points (X,Y,Z Montgomery limbs, 24 words each) -> affine Montgomery, ONE shared inversion
template <class C, int K> HD_INLINE void te_to_affine_shared(typename C::F* ax, typename C::F* ay, const uint32_t* const* p) {
  typedef typename C::F F;
  F X[K], Y[K], Z[K], pre[K];
  for (int k = 0; k < K; k++) for (int i = 0; i < 8; i++) { X[k].v[i] = p[k][i]; Y[k].v[i] = p[k][8 + i]; Z[k].v[i] = p[k][16 + i]; }
  pre[0] = Z[0];
  for (int k = 1; k < K; k++) pre[k] = pre[k - 1] * Z[k];
  F acc = inv(pre[K - 1]);
  for (int k = K - 1; k >= 0; k--) {
    F zi = k ? acc * pre[k - 1] : acc;
    if (k) acc = acc * Z[k];
    ax[k] = X[k] * zi; ay[k] = Y[k] * zi;
  }
}
template <class C> HD_INLINE void store_affine_bytes(uint8_t* out, const typename C::F& x, const typename C::F& y) {
  uint32_t rx[8], ry[8];
  from_mont<typename C::Fq>(rx, x); from_mont<typename C::Fq>(ry, y);
  store_le<8>(out, rx); store_le<8>(out + 32, ry);
}
template <class C> HD_INLINE void load_scalar_bytes_mod_r(uint32_t* k, const uint8_t* p) {   // alignment-free variant
  uint32_t raw[8];
  load_le<8>(raw, p);
  from_mont<typename C::Fr>(k, to_mont<typename C::Fr>(raw));
}

// ietf::Prover::prove, second half (A.9): Y = sk*G, kG, kI already computed (projective)
template <class S>
HD_INLINE void ietf_prove_finish_item(uint8_t* out_c, uint8_t* out_s, const uint8_t* sk_bytes, const uint8_t* k_bytes, const uint8_t* input,
                                      const uint8_t* output, const uint32_t* y_xyz, const uint32_t* kg_xyz, const uint32_t* ki_xyz,
                                      const uint8_t* ad, uint32_t adlen) {
  typedef typename S::C C;
  typename C::F ax[3], ay[3];
  const uint32_t* pp[3] = {y_xyz, kg_xyz, ki_xyz};
  te_to_affine_shared<C, 3>(ax, ay, pp);
  uint8_t enc[5][32];
  ark_encode_point_mont<C>(enc[0], ax[0], ay[0]);
  ark_encode_point_bytes<C>(enc[1], input);
  ark_encode_point_bytes<C>(enc[2], output);
  ark_encode_point_mont<C>(enc[3], ax[1], ay[1]);
  ark_encode_point_mont<C>(enc[4], ax[2], ay[2]);
  uint32_t c[8], sk[8], k[8], sres[8];
  suite_challenge<S>(c, enc, ad, adlen);
  load_scalar_bytes_mod_r<C>(sk, sk_bytes);
  load_le<8>(k, k_bytes);
  scalar_muladd<C>(sres, k, c, sk);
  store_le<8>(out_c, c); store_le<8>(out_s, sres);
}

// Suite::nonce for the rfc8032-style suites, from ABI bytes
template <class S> HD_INLINE void te_nonce_item(uint8_t* out_k, const uint8_t* sk_bytes, const uint8_t* input) {
  typedef typename S::C C;
  uint32_t sk[8], k[8];
  uint8_t enc[32];
  load_scalar_bytes_mod_r<C>(sk, sk_bytes);
  ark_encode_point_bytes<C>(enc, input);
  suite_nonce_8032<S>(k, sk, enc);
  store_le<8>(out_k, k);
}

// pedersen::Prover::prove, first part (A.10): blinding b, nonces k = nonce(sk, I), kb = nonce(b, I)
template <class S> HD_INLINE void pedersen_prove_prep_item(uint8_t* out_b, uint8_t* out_k, uint8_t* out_kb, const uint8_t* sk_bytes,
                                                           const uint8_t* input, const uint8_t* ad, uint32_t adlen) {
  typedef typename S::C C;
  uint32_t sk[8], b[8], k[8], kb[8];
  uint8_t enc[32];
  load_scalar_bytes_mod_r<C>(sk, sk_bytes);
  ark_encode_point_bytes<C>(enc, input);
  suite_blinding<S>(b, sk, enc, ad, adlen);
  suite_nonce_8032<S>(k, sk, enc);
  suite_nonce_8032<S>(kb, b, enc);
  store_le<8>(out_b, b); store_le<8>(out_k, k); store_le<8>(out_kb, kb);
}
// second part: Yb, R, Ok (projective) -> proof bytes (3 x 64 affine || s || sb)
template <class S> HD_INLINE void pedersen_prove_finish_item(uint8_t* proof, const uint8_t* sk_bytes, const uint8_t* b_bytes, const uint8_t* k_bytes,
                                                             const uint8_t* kb_bytes, const uint8_t* input, const uint8_t* output,
                                                             const uint32_t* yb_xyz, const uint32_t* r_xyz, const uint32_t* ok_xyz,
                                                             const uint8_t* ad, uint32_t adlen) {
  typedef typename S::C C;
  typename C::F ax[3], ay[3];
  const uint32_t* pp[3] = {yb_xyz, r_xyz, ok_xyz};
  te_to_affine_shared<C, 3>(ax, ay, pp);
  uint8_t enc[5][32];
  ark_encode_point_mont<C>(enc[0], ax[0], ay[0]);
  ark_encode_point_bytes<C>(enc[1], input);
  ark_encode_point_bytes<C>(enc[2], output);
  ark_encode_point_mont<C>(enc[3], ax[1], ay[1]);
  ark_encode_point_mont<C>(enc[4], ax[2], ay[2]);
  for (int j = 0; j < 3; j++) store_affine_bytes<C>(proof + 64 * j, ax[j], ay[j]);
  uint32_t c[8], sk[8], b[8], k[8], kb[8], r[8];
  suite_challenge<S>(c, enc, ad, adlen);
  load_scalar_bytes_mod_r<C>(sk, sk_bytes);
  load_le<8>(b, b_bytes); load_le<8>(k, k_bytes); load_le<8>(kb, kb_bytes);
  scalar_muladd<C>(r, k, c, sk); store_le<8>(proof + 192, r);
  scalar_muladd<C>(r, kb, c, b); store_le<8>(proof + 224, r);
}
// pedersen::Verifier::verify, first part: c = challenge(Yb, I, O, R, Ok, ad) from the proof bytes
template <class S> HD_INLINE void pedersen_verify_prep_item(uint8_t* out_c, const uint8_t* input, const uint8_t* output, const uint8_t* proof,
                                                            const uint8_t* ad, uint32_t adlen) {
  typedef typename S::C C;
  uint8_t enc[5][32];
  ark_encode_point_bytes<C>(enc[0], proof);
  ark_encode_point_bytes<C>(enc[1], input);
  ark_encode_point_bytes<C>(enc[2], output);
  ark_encode_point_bytes<C>(enc[3], proof + 64);
  ark_encode_point_bytes<C>(enc[4], proof + 128);
  uint32_t c[8];
  suite_challenge<S>(c, enc, ad, adlen);
  store_le<8>(out_c, c);
}
// is the projective point (X,Y,Z) equal to the affine point given as ABI bytes?  (also validates the affine point)
template <class C> HD_INLINE bool te_proj_equals_affine_bytes(const uint32_t* xyz, const uint8_t* aff) {
  typedef typename C::F F;
  uint32_t rx[8], ry[8];
  load_le<8>(rx, aff); load_le<8>(ry, aff + 32);
  bool ok = is_canonical<typename C::Fq>(rx) & is_canonical<typename C::Fq>(ry);
  F x = to_mont<typename C::Fq>(rx), y = to_mont<typename C::Fq>(ry), X, Y, Z;
  ok &= te_on_curve<C>(x, y);
  for (int i = 0; i < 8; i++) { X.v[i] = xyz[i]; Y.v[i] = xyz[8 + i]; Z.v[i] = xyz[16 + i]; }
  return ok & !Z.is_zero() & (x * Z == X) & (y * Z == Y);
}

}  // namespace vrfs
