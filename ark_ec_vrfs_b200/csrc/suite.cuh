// Suite procedures (K5/K6/K9-K11 glue of SURVEY.md 2.5): the device counterpart of the `Suite` trait and
// `utils::{challenge_rfc_9381, nonce_rfc_8032, point_to_hash_rfc_9381}` + `codec::ArkworksCodec`, all
// named at /root/reference/src/lib.rs:13-17 and specified in SURVEY.md Appendix A (A.2, A.6-A.10).
#pragma once
#include "lincomb.cuh"
#include "sha2.cuh"

namespace vrfs {

// per-item Result<(), Error> of the verifiers (include/vrfs_b200.h: vrfs_item_status)
enum { ITEM_OK = 0, ITEM_VERIFICATION_FAILURE = 1, ITEM_INVALID_DATA = 2 };

struct BandSuite {
  typedef BandCurve C;
  typedef Sha512 H;
  static constexpr int CLEN = 32, HLEN = 64, ID_LEN = 25, ENC_LEN = 32;
  static constexpr bool SEC1 = false, RFC6979 = false, ARK_SW = false;
  static constexpr int ID = 0;
  static HD_INLINE uint8_t id(int i) { constexpr char s[] = "Bandersnatch_SHA-512_ELL2"; return (uint8_t)s[i]; }
};
struct EdSuite {
  typedef EdCurve C;
  typedef Sha512 H;
  static constexpr int CLEN = 16, HLEN = 64, ID_LEN = 19, ENC_LEN = 32;
  static constexpr bool SEC1 = false, RFC6979 = false, ARK_SW = false;
  static constexpr int ID = 1;
  static HD_INLINE uint8_t id(int i) { constexpr char s[] = "Ed25519_SHA-512_TAI"; return (uint8_t)s[i]; }
};
struct P256Suite {          // RFC 9381 ECVRF-P256-SHA256-TAI, suite string 0x01
  typedef P256Curve C;
  typedef Sha256 H;
  static constexpr int CLEN = 16, HLEN = 32, ID_LEN = 1, ENC_LEN = 33;
  static constexpr bool SEC1 = true, RFC6979 = true, ARK_SW = false;
  static constexpr int ID = 2;
  static HD_INLINE uint8_t id(int) { return 0x01; }
};

// SURVEY 8(f)4 suites ([RECALL] suite strings, CHALLENGE_LEN, TAI, arkworks codec; blinding bases are placeholders - unpinned)
struct BandSwSuite {        // arkworks codec on a short-Weierstrass curve: 32-byte LE x + a flag byte (ARK_SW)
  typedef BandSwCurve C;
  typedef Sha512 H;
  static constexpr int CLEN = 32, HLEN = 64, ID_LEN = 27, ENC_LEN = 33;
  static constexpr bool SEC1 = false, RFC6979 = false, ARK_SW = true;
  static constexpr int ID = 3;
  static HD_INLINE uint8_t id(int i) { constexpr char s[] = "Bandersnatch_SW_SHA-512_TAI"; return (uint8_t)s[i]; }
};
struct JubSuite {
  typedef JubCurve C;
  typedef Sha512 H;
  static constexpr int CLEN = 32, HLEN = 64, ID_LEN = 18, ENC_LEN = 32;
  static constexpr bool SEC1 = false, RFC6979 = false, ARK_SW = false;
  static constexpr int ID = 4;
  static HD_INLINE uint8_t id(int i) { constexpr char s[] = "JubJub_SHA-512_TAI"; return (uint8_t)s[i]; }
};
struct BjjSuite {
  typedef BjjCurve C;
  typedef Sha512 H;
  static constexpr int CLEN = 32, HLEN = 64, ID_LEN = 22, ENC_LEN = 32;
  static constexpr bool SEC1 = false, RFC6979 = false, ARK_SW = false;
  static constexpr int ID = 5;
  static HD_INLINE uint8_t id(int i) { constexpr char s[] = "BabyJubJub_SHA-512_TAI"; return (uint8_t)s[i]; }
};

template <class S> HD_INLINE void put_suite_id(typename S::H& h) { for (int i = 0; i < S::ID_LEN; i++) h.put(S::id(i)); }

// // codec::Codec::point_encode (A.2) from canonical limbs.
//   ArkworksCodec (TE): 32-byte LE y, bit 255 set iff x > (p-1)/2.
//   Sec1Codec (SW):     (0x02 | (y & 1)) || 32-byte BE x.
template <class S> HD_INLINE void encode_point(uint8_t* out, const uint32_t* x, const uint32_t* y) {
  if (S::SEC1) {
    out[0] = (uint8_t)(2u | (y[0] & 1u));
    store_be<8>(out + 1, x);
  } else if (S::ARK_SW) {     // arkworks short-Weierstrass compressed form: x LE, then the flag byte (bit 7: y > (p-1)/2)
    uint32_t h[8], t[8];
    for (int i = 0; i < 8; i++) h[i] = S::C::Fq::pm1h(i);
    const bool high = MontChains<8>::sub(t, h, y) != 0;
    store_le<8>(out, x);
    out[S::ENC_LEN - 1] = high ? 0x80 : 0x00;
  } else {
    uint32_t h[8], t[8];
    for (int i = 0; i < 8; i++) h[i] = S::C::Fq::pm1h(i);
    bool high = MontChains<8>::sub(t, h, x) != 0;
    store_le<8>(out, y);
    if (high) out[31] |= 0x80;
  }
}
template <class S> HD_INLINE void encode_point_mont(uint8_t* out, const typename S::C::F& x, const typename S::C::F& y) {
  uint32_t rx[8], ry[8];
  from_mont<typename S::C::Fq>(rx, x); from_mont<typename S::C::Fq>(ry, y);
  encode_point<S>(out, rx, ry);
}
// encode straight from the ABI's affine bytes (x||y LE canonical, already validated)
template <class S> HD_INLINE void encode_point_bytes(uint8_t* out, const uint8_t* p) {
  uint32_t rx[8], ry[8];
  load_le<8>(rx, p); load_le<8>(ry, p + 32);
  encode_point<S>(out, rx, ry);
}
// codec scalar_encode: 32 bytes, LE (arkworks) or BE (SEC1)
template <class S> HD_INLINE void encode_scalar(uint8_t* out, const uint32_t* k) {
  if (S::SEC1) store_be<8>(out, k); else store_le<8>(out, k);
}
HD_INLINE bool bytes_all_zero(const uint8_t* p, int n) { uint32_t o = 0; for (int i = 0; i < n; i++) o |= p[i]; return o == 0; }

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
template <class S> HD_INLINE void suite_challenge(uint32_t* c, const uint8_t (*enc)[S::ENC_LEN], const uint8_t* ad, uint32_t adlen) {
  typename S::H h; h.init();
  put_suite_id<S>(h); h.put(0x02);
  for (int p = 0; p < 5; p++) h.update(enc[p], S::ENC_LEN);
  h.update(ad, adlen); h.put(0x00);
  uint8_t dig[S::HLEN]; h.final(dig);
  hash_to_scalar<typename S::C>(c, dig, S::CLEN, true);
}
// point_to_hash_rfc_9381 (A.8)
template <class S> HD_INLINE void suite_point_to_hash(uint8_t* out, const uint8_t* enc) {
  typename S::H h; h.init();
  put_suite_id<S>(h); h.put(0x03); h.update(enc, S::ENC_LEN); h.put(0x00); h.final(out);
}
// nonce_rfc_8032 (A.6): k = LE(H(H(enc_sc(sk))[32..64] || enc_pt(I))) mod r
template <class S> HD_INLINE void suite_nonce_8032(uint32_t* k, const uint32_t* sk, const uint8_t* enc_input) {
  uint8_t e[32], d[64];
  encode_scalar<S>(e, sk);
  typename S::H h; h.init(); h.update(e, 32); h.final(d);
  h.init(); h.update(d + 32, 32); h.update(enc_input, S::ENC_LEN); h.final(d);
  hash_to_scalar<typename S::C>(k, d, 64, false);
}
// nonce_rfc_6979 (A.6; RFC 6979 3.2 with HMAC-SHA-256): x = BE32(sk), h1 = SHA-256(enc_pt(I)) reduced mod n
template <class S> HD_INLINE void suite_nonce_6979(uint32_t* k, const uint32_t* sk, const uint8_t* enc_input) {
  typedef typename S::C C;
  uint8_t h1[32], xb[32], hb[32], V[32], K[32];
  Sha256 h; h.init(); h.update(enc_input, S::ENC_LEN); h.final(h1);
  uint32_t hr[8];
  hash_to_scalar<C>(hr, h1, 32, true);           // bits2octets
  store_be<8>(hb, hr); store_be<8>(xb, sk);
  for (int i = 0; i < 32; i++) { V[i] = 0x01; K[i] = 0x00; }
  HmacSha256 m;
  for (int sep = 0; sep < 2; sep++) {
    uint8_t sb = (uint8_t)sep;
    m.init(K, 32); m.update(V, 32); m.update(&sb, 1); m.update(xb, 32); m.update(hb, 32); m.final(K);
    m.init(K, 32); m.update(V, 32); m.final(V);
  }
  for (;;) {
    m.init(K, 32); m.update(V, 32); m.final(V);
    uint32_t cand[8], nmod[8], t[8];
    load_be<8>(cand, V);
    for (int i = 0; i < 8; i++) nmod[i] = C::Fr::mod(i);
    uint32_t nz = 0;
    for (int i = 0; i < 8; i++) nz |= cand[i];
    if (nz != 0 && MontChains<8>::sub(t, cand, nmod) != 0) { for (int i = 0; i < 8; i++) k[i] = cand[i]; return; }
    uint8_t zero = 0;
    m.init(K, 32); m.update(V, 32); m.update(&zero, 1); m.final(K);
    m.init(K, 32); m.update(V, 32); m.final(V);
  }
}
template <class S> HD_INLINE void suite_nonce(uint32_t* k, const uint32_t* sk, const uint8_t* enc_input) {
  if constexpr (S::RFC6979) suite_nonce_6979<S>(k, sk, enc_input); else suite_nonce_8032<S>(k, sk, enc_input);
}
// Pedersen blinding (A.10): b = BE(H(suite || 0xCC || enc_sc(sk) || enc_pt(I) || ad || 0x00)) mod r
template <class S> HD_INLINE void suite_blinding(uint32_t* b, const uint32_t* sk, const uint8_t* enc_input, const uint8_t* ad, uint32_t adlen) {
  uint8_t e[32], d[64];
  encode_scalar<S>(e, sk);
  typename S::H h; h.init();
  put_suite_id<S>(h); h.put(0xCC); h.update(e, 32); h.update(enc_input, S::ENC_LEN); h.update(ad, adlen); h.put(0x00); h.final(d);
  hash_to_scalar<typename S::C>(b, d, S::HLEN, true);
}
// s = k + c*x mod r (canonical limbs in and out)
template <class C> HD_INLINE void scalar_muladd(uint32_t* s, const uint32_t* k, const uint32_t* c, const uint32_t* x) {
  typedef typename C::Fr Fr;
  Fp<Fr> km = to_mont<Fr>(k), cm = to_mont<Fr>(c), xm = to_mont<Fr>(x);
  from_mont<Fr>(s, km + cm * xm);
}

// two projective points (X,Y,Z Montgomery limbs) -> affine with one shared inversion
template <class C> HD_INLINE void two_to_affine(typename C::F* ax, typename C::F* ay, const uint32_t* p0, const uint32_t* p1) {
  typedef typename C::F F;
  F X0, Y0, Z0, X1, Y1, Z1;
  for (int i = 0; i < 8; i++) { X0.v[i] = p0[i]; Y0.v[i] = p0[8 + i]; Z0.v[i] = p0[16 + i]; X1.v[i] = p1[i]; Y1.v[i] = p1[8 + i]; Z1.v[i] = p1[16 + i]; }
  F ti = inv(Z0 * Z1);
  F z0i = ti * Z1, z1i = ti * Z0;
  ax[0] = X0 * z0i; ay[0] = Y0 * z0i; ax[1] = X1 * z1i; ay[1] = Y1 * z1i;
}

// ietf::Verifier::verify, second half (A.9): U, V already computed; accept iff challenge(Y,I,O,U,V,ad) == c.
// zinv = 1 / (Z_U * Z_V), shared by the two points (and itself obtained from a batched inversion, see
// ietf_verify_finish_batched).
template <class S>
HD_INLINE bool ietf_verify_finish_with_zinv(const uint8_t* pk, const uint8_t* input, const uint8_t* output, const uint8_t* c_bytes,
                                            const uint32_t* u_xyz, const uint32_t* v_xyz, const typename S::C::F& zinv,
                                            const uint8_t* ad, uint32_t adlen) {
  typedef typename S::C C;
  typedef typename C::F F;
  F XU, YU, ZU, XV, YV, ZV;
  for (int i = 0; i < 8; i++) { XU.v[i] = u_xyz[i]; YU.v[i] = u_xyz[8 + i]; ZU.v[i] = u_xyz[16 + i]; XV.v[i] = v_xyz[i]; YV.v[i] = v_xyz[8 + i]; ZV.v[i] = v_xyz[16 + i]; }
  if (!C::IS_TE) {   // the identity has no SEC1-compressed encoding: Error::InvalidData / VerificationFailure
    if (ZU.is_zero() || ZV.is_zero() || bytes_all_zero(pk, 64) || bytes_all_zero(input, 64) || bytes_all_zero(output, 64)) return false;
  }
  F zui = zinv * ZV, zvi = zinv * ZU;
  uint8_t enc[5][S::ENC_LEN];
  encode_point_bytes<S>(enc[0], pk);
  encode_point_bytes<S>(enc[1], input);
  encode_point_bytes<S>(enc[2], output);
  encode_point_mont<S>(enc[3], XU * zui, YU * zui);
  encode_point_mont<S>(enc[4], XV * zvi, YV * zvi);
  uint32_t c2[8], c[8];
  suite_challenge<S>(c2, enc, ad, adlen);
  hash_to_scalar<C>(c, c_bytes, 32, false);
  uint32_t diff = 0;
  for (int i = 0; i < 8; i++) diff |= c[i] ^ c2[i];
  return diff == 0;
}
template <class C> HD_INLINE typename C::F zz_product(const uint32_t* u_xyz, const uint32_t* v_xyz) {
  typename C::F ZU, ZV;
  for (int i = 0; i < 8; i++) { ZU.v[i] = u_xyz[16 + i]; ZV.v[i] = v_xyz[16 + i]; }
  return ZU * ZV;
}
template <class S>
HD_INLINE bool ietf_verify_finish_item(const uint8_t* pk, const uint8_t* input, const uint8_t* output, const uint8_t* c_bytes,
                                       const uint32_t* u_xyz, const uint32_t* v_xyz, const uint8_t* ad, uint32_t adlen) {
  typename S::C::F zinv = inv(zz_product<typename S::C>(u_xyz, v_xyz));
  return ietf_verify_finish_with_zinv<S>(pk, input, output, c_bytes, u_xyz, v_xyz, zinv, ad, adlen);
}
// K items per thread share ONE field inversion (Montgomery's trick): item k of this thread is `first + k * stride`.
// A zero Z (only possible for invalid / identity inputs) is replaced by 1 in the product chain so that it cannot
// poison the other items; that item's own zinv is then wrong-but-unused (it fails the Z checks / validity flag).
template <class S, int K>
HD_INLINE void ietf_verify_finish_batched(uint32_t n, uint32_t first, uint32_t stride, const uint8_t* pk, const uint8_t* input, const uint8_t* output,
                                          const uint8_t* c, const uint32_t* u_xyz, const uint32_t* v_xyz, const uint8_t* ad, const uint64_t* ad_off,
                                          const uint8_t* valid, uint8_t* out_ok, uint8_t* out_status = nullptr) {
  typedef typename S::C C;
  typedef typename C::F F;
  F z[K], pre[K];
  int cnt = 0;
  for (int k = 0; k < K; k++) {
    uint32_t i = first + (uint32_t)k * stride;
    if (i >= n) break;
    F t = zz_product<C>(u_xyz + (size_t)24 * i, v_xyz + (size_t)24 * i);
    z[k] = select(t.is_zero(), F::one(), t);
    pre[k] = k ? pre[k - 1] * z[k] : z[k];
    cnt = k + 1;
  }
  if (cnt == 0) return;
  F acc = inv(pre[cnt - 1]);
  for (int k = K - 1; k >= 0; k--) {
    if (k >= cnt) continue;
    uint32_t i = first + (uint32_t)k * stride;
    F zinv = k ? acc * pre[k - 1] : acc;
    if (k) acc = acc * z[k];
    const uint8_t* a = ad ? ad + ad_off[i] : nullptr;
    uint32_t alen = ad ? (uint32_t)(ad_off[i + 1] - ad_off[i]) : 0u;
    bool ok = ietf_verify_finish_with_zinv<S>(pk + (size_t)64 * i, input + (size_t)64 * i, output + (size_t)64 * i, c + (size_t)32 * i,
                                              u_xyz + (size_t)24 * i, v_xyz + (size_t)24 * i, zinv, a, alen);
    // TE with Z = 0 cannot happen for on-curve inputs; guard anyway so a forged zero never verifies
    F t = zz_product<C>(u_xyz + (size_t)24 * i, v_xyz + (size_t)24 * i);
    const bool good = ok && valid[i] && !t.is_zero();
    out_ok[i] = (uint8_t)good;
    if (out_status) {     // Result<(), Error>: values no typed Public / Input / Output can hold are InvalidData, the rest VerificationFailure
      bool bad_input = !valid[i];
      if (!C::IS_TE) bad_input |= bytes_all_zero(pk + (size_t)64 * i, 64) || bytes_all_zero(input + (size_t)64 * i, 64) || bytes_all_zero(output + (size_t)64 * i, 64);
      out_status[i] = (uint8_t)(good ? ITEM_OK : bad_input ? ITEM_INVALID_DATA : ITEM_VERIFICATION_FAILURE);
    }
  }
}

// K projective points (X,Y,Z Montgomery limbs, 24 words each) -> affine Montgomery, ONE shared inversion.  A point with
// Z = 0 (the short-Weierstrass identity) comes out as (0,0) and does not disturb the others.  `zinv` (optional): the inverse
// of the product of the non-zero Z's, precomputed for many items at once by zinv_batched (one inversion per 8 items).
template <class C, int K> HD_INLINE void to_affine_shared(typename C::F* ax, typename C::F* ay, const uint32_t* const* p, const uint32_t* zinv = nullptr) {
  typedef typename C::F F;
  F X[K], Y[K], Z[K], pre[K];
  bool zero[K];
  for (int k = 0; k < K; k++) {
    for (int i = 0; i < 8; i++) { X[k].v[i] = p[k][i]; Y[k].v[i] = p[k][8 + i]; Z[k].v[i] = p[k][16 + i]; }
    zero[k] = Z[k].is_zero();
    Z[k] = select(zero[k], F::one(), Z[k]);
  }
  pre[0] = Z[0];
  for (int k = 1; k < K; k++) pre[k] = pre[k - 1] * Z[k];
  F acc;
  if (zinv) { for (int i = 0; i < 8; i++) acc.v[i] = zinv[i]; } else acc = inv(pre[K - 1]);
  for (int k = K - 1; k >= 0; k--) {
    F zi = k ? acc * pre[k - 1] : acc;
    if (k) acc = acc * Z[k];
    zi = select(zero[k], F::zero(), zi);
    ax[k] = X[k] * zi; ay[k] = Y[k] * zi;
  }
}
// zinv[i] = 1 / prod_j Z_j(i) over the NP projective points of item i (zero Z's skipped), for the K items
// first, first + stride, ... of this thread with ONE field inversion (Montgomery's trick)
template <class C, int NP, int K>
HD_INLINE void zinv_batched(uint32_t n, uint32_t first, uint32_t stride, const uint32_t* const* pts, uint32_t* zinv) {
  typedef typename C::F F;
  F z[K], pre[K];
  int cnt = 0;
  for (int k = 0; k < K; k++) {
    uint32_t i = first + (uint32_t)k * stride;
    if (i >= n) break;
    F t = F::one();
    for (int j = 0; j < NP; j++) {
      F Zj;
      for (int w = 0; w < 8; w++) Zj.v[w] = pts[j][(size_t)24 * i + 16 + w];
      Zj = select(Zj.is_zero(), F::one(), Zj);
      t = j ? t * Zj : Zj;
    }
    z[k] = t;
    pre[k] = k ? pre[k - 1] * t : t;
    cnt = k + 1;
  }
  if (cnt == 0) return;
  F acc = inv(pre[cnt - 1]);
  for (int k = K - 1; k >= 0; k--) {
    if (k >= cnt) continue;
    uint32_t i = first + (uint32_t)k * stride;
    F zi = k ? acc * pre[k - 1] : acc;
    if (k) acc = acc * z[k];
    for (int w = 0; w < 8; w++) zinv[(size_t)8 * i + w] = zi.v[w];
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
                                      const uint8_t* ad, uint32_t adlen, const uint32_t* zinv = nullptr) {
  typedef typename S::C C;
  typename C::F ax[3], ay[3];
  const uint32_t* pp[3] = {y_xyz, kg_xyz, ki_xyz};
  to_affine_shared<C, 3>(ax, ay, pp, zinv);
  uint8_t enc[5][S::ENC_LEN];
  encode_point_mont<S>(enc[0], ax[0], ay[0]);
  encode_point_bytes<S>(enc[1], input);
  encode_point_bytes<S>(enc[2], output);
  encode_point_mont<S>(enc[3], ax[1], ay[1]);
  encode_point_mont<S>(enc[4], ax[2], ay[2]);
  uint32_t c[8], sk[8], k[8], sres[8];
  suite_challenge<S>(c, enc, ad, adlen);
  load_scalar_bytes_mod_r<C>(sk, sk_bytes);
  load_le<8>(k, k_bytes);
  scalar_muladd<C>(sres, k, c, sk);
  store_le<8>(out_c, c); store_le<8>(out_s, sres);
}

// Suite::nonce from ABI bytes
template <class S> HD_INLINE void nonce_item(uint8_t* out_k, const uint8_t* sk_bytes, const uint8_t* input) {
  typedef typename S::C C;
  uint32_t sk[8], k[8];
  uint8_t enc[S::ENC_LEN];
  load_scalar_bytes_mod_r<C>(sk, sk_bytes);
  encode_point_bytes<S>(enc, input);
  suite_nonce<S>(k, sk, enc);
  store_le<8>(out_k, k);
}

// pedersen::Prover::prove, first part (A.10): blinding b, nonces k = nonce(sk, I), kb = nonce(b, I)
template <class S> HD_INLINE void pedersen_prove_prep_item(uint8_t* out_b, uint8_t* out_k, uint8_t* out_kb, const uint8_t* sk_bytes,
                                                           const uint8_t* input, const uint8_t* ad, uint32_t adlen) {
  typedef typename S::C C;
  uint32_t sk[8], b[8], k[8], kb[8];
  uint8_t enc[S::ENC_LEN];
  load_scalar_bytes_mod_r<C>(sk, sk_bytes);
  encode_point_bytes<S>(enc, input);
  suite_blinding<S>(b, sk, enc, ad, adlen);
  suite_nonce<S>(k, sk, enc);
  suite_nonce<S>(kb, b, enc);
  store_le<8>(out_b, b); store_le<8>(out_k, k); store_le<8>(out_kb, kb);
}
// second part: Yb, R, Ok (projective) -> proof bytes (3 x 64 affine || s || sb)
template <class S> HD_INLINE void pedersen_prove_finish_item(uint8_t* proof, const uint8_t* sk_bytes, const uint8_t* b_bytes, const uint8_t* k_bytes,
                                                             const uint8_t* kb_bytes, const uint8_t* input, const uint8_t* output,
                                                             const uint32_t* yb_xyz, const uint32_t* r_xyz, const uint32_t* ok_xyz,
                                                             const uint8_t* ad, uint32_t adlen, const uint32_t* zinv = nullptr) {
  typedef typename S::C C;
  typename C::F ax[3], ay[3];
  const uint32_t* pp[3] = {yb_xyz, r_xyz, ok_xyz};
  to_affine_shared<C, 3>(ax, ay, pp, zinv);
  uint8_t enc[5][S::ENC_LEN];
  encode_point_mont<S>(enc[0], ax[0], ay[0]);
  encode_point_bytes<S>(enc[1], input);
  encode_point_bytes<S>(enc[2], output);
  encode_point_mont<S>(enc[3], ax[1], ay[1]);
  encode_point_mont<S>(enc[4], ax[2], ay[2]);
  for (int j = 0; j < 3; j++) store_affine_bytes<C>(proof + 64 * j, ax[j], ay[j]);
  uint32_t c[8], sk[8], b[8], k[8], kb[8], r[8];
  suite_challenge<S>(c, enc, ad, adlen);
  load_scalar_bytes_mod_r<C>(sk, sk_bytes);
  load_le<8>(b, b_bytes); load_le<8>(k, k_bytes); load_le<8>(kb, kb_bytes);
  scalar_muladd<C>(r, k, c, sk); store_le<8>(proof + 192, r);
  scalar_muladd<C>(r, kb, c, b); store_le<8>(proof + 224, r);
}
// pedersen::Verifier::verify, first part: c = challenge(Yb, I, O, R, Ok, ad) from the proof bytes
// returns false when a point is the short-Weierstrass identity (not encodable -> the reference rejects)
template <class S> HD_INLINE bool pedersen_verify_prep_item(uint8_t* out_c, const uint8_t* input, const uint8_t* output, const uint8_t* proof,
                                                            const uint8_t* ad, uint32_t adlen) {
  typedef typename S::C C;
  bool ok = true;
  if (!C::IS_TE) ok = !(bytes_all_zero(input, 64) || bytes_all_zero(output, 64) || bytes_all_zero(proof, 64) || bytes_all_zero(proof + 64, 64) || bytes_all_zero(proof + 128, 64));
  uint8_t enc[5][S::ENC_LEN];
  encode_point_bytes<S>(enc[0], proof);
  encode_point_bytes<S>(enc[1], input);
  encode_point_bytes<S>(enc[2], output);
  encode_point_bytes<S>(enc[3], proof + 64);
  encode_point_bytes<S>(enc[4], proof + 128);
  uint32_t c[8];
  suite_challenge<S>(c, enc, ad, adlen);
  store_le<8>(out_c, c);
  return ok;
}
// is the projective point (X,Y,Z) equal to the affine point given as ABI bytes?  Also validates the affine point:
// 0 = the bytes are no curve point (non-canonical / off the curve), 1 = a curve point different from (X:Y:Z), 2 = equal
template <class C> HD_INLINE int proj_vs_affine_bytes(const uint32_t* xyz, const uint8_t* aff) {
  typedef typename C::F F;
  uint32_t rx[8], ry[8];
  load_le<8>(rx, aff); load_le<8>(ry, aff + 32);
  bool ok = is_canonical<typename C::Fq>(rx) & is_canonical<typename C::Fq>(ry);
  F x = to_mont<typename C::Fq>(rx), y = to_mont<typename C::Fq>(ry), X, Y, Z;
  ok &= Grp<C>::on_curve(x, y);
  for (int i = 0; i < 8; i++) { X.v[i] = xyz[i]; Y.v[i] = xyz[8 + i]; Z.v[i] = xyz[16 + i]; }
  if (!ok) return 0;
  return (!Z.is_zero() & (x * Z == X) & (y * Z == Y)) ? 2 : 1;
}

}  // namespace vrfs
