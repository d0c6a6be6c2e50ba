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

}  // namespace vrfs
