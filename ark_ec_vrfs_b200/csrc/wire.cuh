// Wire formats: the device side of `CanonicalDeserialize` (Compress::Yes, Validate::Yes) for
// `Public` / `Output` / `ietf::Proof` and of `CanonicalSerialize` for a signature - names re-exported at
// /root/reference/src/lib.rs:13-17 (`Public, Output, ietf, codec`).
//   point  : codec point_decode (A.2) + on-curve + prime-order-subgroup membership.  The reference's subgroup test is
//            ark-ec's default `mul_bigint(r).is_zero()`; any test deciding the same predicate gives the same verdict.
//   proof  : c = CHALLENGE_LEN bytes in codec byte order (reduced mod r), s = 32 bytes, rejected when >= r   (A.9)
//   signature = point_encode(Output) || c || s      (Bandersnatch 96 B, Ed25519 80 B, secp256r1 81 B = RFC 9381 pi_string)
// secp256r1 caveat (parity unpinned): what is implemented for that suite is the CODEC's form - SEC1 points and
// big-endian scalars, i.e. RFC 9381's pi_string, pinned by the RFC's Examples 10-11.  Whether the crate's derived
// CanonicalSerialize of `Public` / `Output` / `pedersen::Proof` uses this or arkworks' native short-Weierstrass form (33-byte
// little-endian x || flags, little-endian s - the form the engine implements for bandersnatch_sw, `ARK_SW`) cannot be settled
// without the crate; a caller that needs the latter for secp256r1 must re-encode.  Bandersnatch and Ed25519 are unaffected.
#pragma once
#include "h2c.cuh"

namespace vrfs {

// ---- prime-order subgroup membership of an affine point that is on the curve -----------------------------------
// Bandersnatch: E(Fq) = Z/2 x Z/2 x Z/r with only (0,1), (0,-1) of the 2-torsion affine, hence <G> = 2*E(Fq), and
// by 2-descent on the Montgomery model B t^2 = s (s - alpha)(s - 1/alpha), s = (1+y)/(1-y):
//     P in 2E  <=>  B*s and B*(s - alpha) are non-zero squares
//              <=>  chi(B (1+y)(1-y)) = chi(B ((1+y) - alpha (1-y)) (1-y)) = 1.
// Two quadratic characters (binary Jacobi symbols, ~20 K ALU instructions each) instead of a 253-bit scalar multiplication;
// checked against [r]P on all four cosets in tests (oracle_subgroup_check_batch restates the reference's test).
template <class C> struct SubgroupCheck;
template <> struct SubgroupCheck<BandCurve> {
  static HD_INLINE bool run(const BandCurve::F& x, const BandCurve::F& y) {
    typedef BandCurve::F F;
    const F one = F::one();
    if (x.is_zero() && y == one) return true;
    const F B = fconst<BlsFr, BandConsts::ELL2_K>(), alpha = fconst<BlsFr, BandConsts::SUBGRP_ALPHA>();
    F p = one + y, m = one - y;
    F s1 = B * p * m, s2 = B * (p - alpha * m) * m;
    // non-zero squares, decided by the binary Jacobi symbol (no multiplier work); zero covers (0,-1) and anything degenerate
    return jacobi(s1) == 1 && jacobi(s2) == 1;
  }
};
// Ed25519 (cofactor 8, cyclic torsion): [L]P == O with the complete a = -1 addition law; L's bits are compile-time
// constants, so the double-and-add branches are warp-uniform.
template <> struct SubgroupCheck<EdCurve> {
  static HD_INLINE bool run(const EdCurve::F& x, const EdCurve::F& y) {
    TEPoint<EdCurve> P, R;
    te_from_affine<EdCurve>(P, x, y);
    te_set_identity(R);
#pragma unroll 1
    for (int i = 252; i >= 0; i--) {
      te_dbl<EdCurve>(&R, &R, true);
      if ((EdFr::mod(i >> 5) >> (i & 31)) & 1u) te_add<EdCurve>(&R, &R, &P);
    }
    return R.X.is_zero() && R.Y == R.Z && !R.Y.is_zero();
  }
};
// [r]P == O by double-and-add over the public bits of r (warp-uniform branches), for the curves added by SURVEY 8(f)4
template <class C> struct SubgroupCheckMulR {
  static HD_INLINE bool run(const typename C::F& x, const typename C::F& y) {
    typedef Grp<C> G;
    typename G::Pt P, Rr;
    G::from_affine(P, x, y);
    G::set_identity(Rr);
#pragma unroll 1
    for (int i = 255; i >= 0; i--) {
      G::dbl(&Rr);
      if ((C::Fr::mod(i >> 5) >> (i & 31)) & 1u) G::add(&Rr, &Rr, &P);
    }
    if constexpr (C::IS_TE) return Rr.X.is_zero() && Rr.Y == Rr.Z && !Rr.Y.is_zero();
    else return Rr.Z.is_zero();
  }
};
template <> struct SubgroupCheck<JubCurve> : SubgroupCheckMulR<JubCurve> {};
template <> struct SubgroupCheck<BjjCurve> : SubgroupCheckMulR<BjjCurve> {};
template <> struct SubgroupCheck<BandSwCurve> : SubgroupCheckMulR<BandSwCurve> {};
template <> struct SubgroupCheck<P256Curve> {           // cofactor 1
  static HD_INLINE bool run(const P256Curve::F&, const P256Curve::F&) { return true; }
};

// Public / Output deserialisation: enc -> affine ABI bytes (x || y LE); false (and zero bytes) when rejected
template <class S> HD_INLINE bool wire_decode_point_checked(uint8_t* out64, const uint8_t* enc) {
  typedef typename S::C C;
  typename C::F x, y;
  bool ok = decode_point<S>(x, y, enc);
  if (ok) ok = SubgroupCheck<C>::run(x, y);
  if (ok) store_affine_bytes<C>(out64, x, y); else for (int j = 0; j < 64; j++) out64[j] = 0;
  return ok;
}

// ietf::Proof deserialisation: c32 / s32 = 32-byte little-endian ABI scalars; false when s is not canonical
template <class S> HD_INLINE bool wire_parse_proof(uint8_t* c32, uint8_t* s32, const uint8_t* proof) {
  for (int j = 0; j < 32; j++) c32[j] = 0;
  for (int j = 0; j < S::CLEN; j++) c32[j] = S::SEC1 ? proof[S::CLEN - 1 - j] : proof[j];
  for (int j = 0; j < 32; j++) s32[j] = S::SEC1 ? proof[S::CLEN + 31 - j] : proof[S::CLEN + j];
  uint32_t raw[8];
  load_le<8>(raw, s32);
  return is_canonical<typename S::C::Fr>(raw);
}
// signature serialisation from the ABI values (affine output, 32-byte LE c and s)
template <class S> HD_INLINE void wire_pack_signature(uint8_t* sig, const uint8_t* output64, const uint8_t* c32, const uint8_t* s32) {
  encode_point_bytes<S>(sig, output64);
  uint8_t* p = sig + S::ENC_LEN;
  for (int j = 0; j < S::CLEN; j++) p[j] = S::SEC1 ? c32[S::CLEN - 1 - j] : c32[j];
  for (int j = 0; j < 32; j++) p[S::CLEN + j] = S::SEC1 ? s32[31 - j] : s32[j];
}

}  // namespace vrfs
