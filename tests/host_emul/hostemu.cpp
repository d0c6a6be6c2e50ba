// TEST INFRASTRUCTURE ONLY: compiles the engine's device headers for the HOST (every PTX block has a
// plain-C twin) so that the field / curve / hashing code can be checked on a box without a GPU.
// Nothing in the product links this file.
#include <stdint.h>
#include <string.h>
#include "../../ark_ec_vrfs_b200/csrc/gen/field_consts.cuh"
using namespace vrfs;

template <class P> static void field_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Fp<P> x, y, r;
  memcpy(x.v, a, 4 * P::N); memcpy(y.v, b, 4 * P::N);
  switch (op) {
    case 0: r = x * y; break;                       // Montgomery product of raw limb arrays
    case 1: r = x + y; break;
    case 2: r = x - y; break;
    case 3: r = to_mont<P>(a); break;               // any value -> Montgomery, reduced
    case 4: from_mont<P>(r.v, x); break;
    case 5: r = inv(x); break;
    case 6: r = to_mont_wide<P>(a, b); break;
    case 7: r = Fp<P>::zero(); r.v[0] = is_square(x); break;
    case 8: r = Fp<P>::zero(); r.v[0] = is_high(x) | (is_odd(x) << 1); break;
    case 9: r = sqr(x); break;                      // dedicated squaring path (SOS reduction)
    case 10: r = Fp<P>::zero(); r.v[0] = (uint32_t)(jacobi(x) + 1); break;   // Jacobi symbol + 1
    default: r = Fp<P>::zero();
  }
  memcpy(out, r.v, 4 * P::N);
}
extern "C" void hostemu_field_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  switch (field) {
    case 0: field_op<BlsFr>(op, a, b, out); break;
    case 1: field_op<BandFr>(op, a, b, out); break;
    case 2: field_op<F25519>(op, a, b, out); break;
    case 3: field_op<EdFr>(op, a, b, out); break;
    case 4: field_op<P256Fp>(op, a, b, out); break;
    case 5: field_op<P256Fr>(op, a, b, out); break;
    case 6: field_op<BlsFq>(op, a, b, out); break;
  }
}

// ---------------------------------------------------------------------------------------------
// per-item engine code on the host: IETF verify (Bandersnatch / Ed25519)
#include <vector>
#include "../../ark_ec_vrfs_b200/csrc/h2c.cuh"

template <class C> static const typename Grp<C>::FixEntry* fixed_table(bool blinding) {
  static std::vector<typename Grp<C>::FixEntry> tabs[2];
  auto& t = tabs[blinding];
  if (t.empty()) {
    t.resize(fix_table_entries<C>());
    for (int w = 0; w < Grp<C>::FIX_WINDOWS; w++) for (int d = 0; d < FIX_ENTRIES; d++)
      fixed_table_entry<C>(t[(size_t)w * FIX_ENTRIES + d], blinding ? C::bx() : C::gx(), blinding ? C::by() : C::gy(), w, d);
  }
  return t.data();
}

template <class S> static void ietf_verify(size_t n, const uint8_t* pk, const uint8_t* in, const uint8_t* out, const uint8_t* c,
                                           const uint8_t* s, const uint8_t* ad, const uint64_t* ad_off, uint8_t* ok) {
  typedef typename S::C C;
  std::vector<typename Grp<C>::Entry> slab(4 * 9);
  std::vector<uint32_t> u(28), v(28), us(24 * n), vs(24 * n);
  std::vector<uint8_t> valids(n);
  for (size_t i = 0; i < n; i++) {
    LincombArgs A = {};
    A.n = (uint32_t)n;
    const uint32_t cbits = S::CLEN < 32 ? 8u * S::CLEN : 0u;   // like ietf_verify_dev: short challenges skip their empty windows
    A.var[0] = {pk, 64, c, 32, 1, cbits};
    A.fix[0] = {s, 32, 0, fixed_table<C>(false)};
    typename Grp<C>::Pt acc;
    bool valid = lincomb_item<C, 1, 1>(A, (uint32_t)i, slab.data(), acc);
    Grp<C>::store_xyz(u.data(), acc);
    A.var[0] = {in, 64, s, 32, 0};
    A.var[1] = {out, 64, c, 32, 1, cbits};
    valid &= lincomb_item<C, 2, 0>(A, (uint32_t)i, slab.data(), acc);
    Grp<C>::store_xyz(v.data(), acc);
    const uint8_t* a = ad ? ad + ad_off[i] : (const uint8_t*)"";
    uint32_t alen = ad ? (uint32_t)(ad_off[i + 1] - ad_off[i]) : 0;
    ok[i] = valid && ietf_verify_finish_item<S>(pk + 64 * i, in + 64 * i, out + 64 * i, c + 32 * i, u.data(), v.data(), a, alen);
    memcpy(&us[24 * i], u.data(), 96); memcpy(&vs[24 * i], v.data(), 96); valids[i] = valid;
  }
  // the batched-inversion form used by the kernel must agree item by item
  std::vector<uint8_t> ok2(n, 7);
  const uint32_t threads = (uint32_t)((n + 2) / 3);
  for (uint32_t t = 0; t < threads; t++)
    ietf_verify_finish_batched<S, 3>((uint32_t)n, t, threads, pk, in, out, c, us.data(), vs.data(), ad, ad_off, valids.data(), ok2.data());
  for (size_t i = 0; i < n; i++) if (ok2[i] != ok[i]) ok[i] = 0xEE;
}
extern "C" void hostemu_ietf_verify(int suite, size_t n, const uint8_t* pk, const uint8_t* in, const uint8_t* out, const uint8_t* c,
                                    const uint8_t* s, const uint8_t* ad, const uint64_t* ad_off, uint8_t* ok) {
  if (suite == 0) ietf_verify<BandSuite>(n, pk, in, out, c, s, ad, ad_off, ok);
  else if (suite == 1) ietf_verify<EdSuite>(n, pk, in, out, c, s, ad, ad_off, ok);
  else ietf_verify<P256Suite>(n, pk, in, out, c, s, ad, ad_off, ok);
}

// debug / unit-test entry: R = k1*P1 [+ k2*P2] [+ f*G], affine canonical out (x||y LE)
template <class C> static int lincomb_dbg(int nv, int nf, const uint8_t* p1, const uint8_t* k1, const uint8_t* p2, const uint8_t* k2,
                                          const uint8_t* f, int neg_mask, uint8_t* out) {
  std::vector<typename Grp<C>::Entry> slab(4 * 9);
  LincombArgs A = {};
  A.n = 1;
  A.var[0] = {p1, 64, k1, 32, (uint32_t)(neg_mask & 1)};
  A.var[1] = {p2, 64, k2, 32, (uint32_t)((neg_mask >> 1) & 1)};
  A.fix[0] = {f, 32, (uint32_t)((neg_mask >> 2) & 1), fixed_table<C>(false)};
  typename Grp<C>::Pt acc; bool ok;
  if (nv == 1 && nf == 0) ok = lincomb_item<C, 1, 0>(A, 0, slab.data(), acc);
  else if (nv == 2 && nf == 0) ok = lincomb_item<C, 2, 0>(A, 0, slab.data(), acc);
  else if (nv == 0 && nf == 1) ok = lincomb_item<C, 0, 1>(A, 0, slab.data(), acc);
  else ok = lincomb_item<C, 1, 1>(A, 0, slab.data(), acc);
  alignas(16) uint32_t xyz[24];
  Grp<C>::store_xyz(xyz, acc);
  typename C::F X, Y, Z;
  memcpy(X.v, xyz, 32); memcpy(Y.v, xyz + 8, 32); memcpy(Z.v, xyz + 16, 32);
  typename C::F zi = inv(Z), x = X * zi, y = Y * zi;
  uint32_t rx[8], ry[8]; from_mont<typename C::Fq>(rx, x); from_mont<typename C::Fq>(ry, y);
  store_le<8>(out, rx); store_le<8>(out + 32, ry);
  return ok;
}
extern "C" int hostemu_lincomb(int suite, int nv, int nf, const uint8_t* p1, const uint8_t* k1, const uint8_t* p2, const uint8_t* k2,
                               const uint8_t* f, int neg_mask, uint8_t* out) {
  return suite == 0 ? lincomb_dbg<BandCurve>(nv, nf, p1, k1, p2, k2, f, neg_mask, out)
       : suite == 1 ? lincomb_dbg<EdCurve>(nv, nf, p1, k1, p2, k2, f, neg_mask, out) : lincomb_dbg<P256Curve>(nv, nf, p1, k1, p2, k2, f, neg_mask, out);
}
extern "C" void hostemu_glv(const uint32_t* k, uint32_t* out /*4+4+2*/) {
  GlvHalf a, b; band_glv_split(&a, &b, k);
  memcpy(out, a.mag, 16); memcpy(out + 4, b.mag, 16); out[8] = a.neg; out[9] = b.neg;
}

// wire formats (csrc/wire.cuh) on the host: checked point decode, proof parse, signature pack
#include "../../ark_ec_vrfs_b200/csrc/wire.cuh"
template <class S> static void wire_decode_checked(size_t n, const uint8_t* enc, uint8_t* out, uint8_t* ok) {
  for (size_t i = 0; i < n; i++) ok[i] = wire_decode_point_checked<S>(out + 64 * i, enc + (size_t)S::ENC_LEN * i);
}
extern "C" void hostemu_decode_checked(int suite, size_t n, const uint8_t* enc, uint8_t* out, uint8_t* ok) {
  if (suite == 0) wire_decode_checked<BandSuite>(n, enc, out, ok);
  else if (suite == 1) wire_decode_checked<EdSuite>(n, enc, out, ok);
  else wire_decode_checked<P256Suite>(n, enc, out, ok);
}
template <class S> static void wire_roundtrip(const uint8_t* out64, const uint8_t* c32, const uint8_t* s32, uint8_t* sig, uint8_t* c_back, uint8_t* s_back, uint8_t* ok) {
  wire_pack_signature<S>(sig, out64, c32, s32);
  *ok = wire_parse_proof<S>(c_back, s_back, sig + S::ENC_LEN);
}
extern "C" void hostemu_wire_roundtrip(int suite, const uint8_t* out64, const uint8_t* c32, const uint8_t* s32, uint8_t* sig, uint8_t* c_back, uint8_t* s_back, uint8_t* ok) {
  if (suite == 0) wire_roundtrip<BandSuite>(out64, c32, s32, sig, c_back, s_back, ok);
  else if (suite == 1) wire_roundtrip<EdSuite>(out64, c32, s32, sig, c_back, s_back, ok);
  else wire_roundtrip<P256Suite>(out64, c32, s32, sig, c_back, s_back, ok);
}

// Suite::data_to_point for Bandersnatch (Elligator2 with the table-based torsion logarithm) and a bare square root
extern "C" void hostemu_band_h2c(const uint8_t* data, uint32_t len, uint8_t* out64) {
  TEPoint<BandCurve> P;
  band_h2c_ell2(P, data, len);
  Fp<BlsFr> zi = inv(P.Z);
  store_affine_bytes<BandCurve>(out64, P.X * zi, P.Y * zi);
}
extern "C" int hostemu_bls_sqrt(const uint32_t* a_mont, uint32_t* out_mont) {
  Fp<BlsFr> a, r; memcpy(a.v, a_mont, 32);
  bool ok = sqrt_ct<BlsFr>(&r, &a);
  memcpy(out_mont, r.v, 32);
  return ok;
}

// BLS12-381 Fq inversion by binary extended Euclid (csrc/msm.cuh), Montgomery in / Montgomery out
#include "../../ark_ec_vrfs_b200/csrc/msm.cuh"
extern "C" void hostemu_fq381_inv(const uint32_t* a_mont, uint32_t* out_mont) {
  Fq381 a; memcpy(a.v, a_mont, 48);
  Fq381 r = fq381_inv(a);
  memcpy(out_mont, r.v, 48);
}
// the word-approximation binary GCD; returns 1 when its own loop finished (no fallback to fq381_inv was needed)
extern "C" int hostemu_fq381_inv_fast(const uint32_t* a_mont, uint32_t* out_mont) {
  Fq381 a, t; memcpy(a.v, a_mont, 48);
  const bool finished = fq381_inv_bingcd(t, a);
  Fq381 r = fq381_inv_fast(a);
  memcpy(out_mont, r.v, 48);
  return finished;
}

// k = k1 + q z^2 (csrc/msm.cuh g1_glv_split, the stateless MSM's scalar halves); canonical limbs in and out
extern "C" void hostemu_g1_glv_split(const uint32_t* k8, uint32_t* k1_4, uint32_t* q_4) { g1_glv_split(k1_4, q_4, k8); }
// a b - c d with one reduction (csrc/msm.cuh fq381_mul_sub2, the Y coordinate of the bucket addition); Montgomery limbs in and out
extern "C" void hostemu_fq381_mul_sub2(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* out_mont) {
  Fq381 A, B, Cc, D; memcpy(A.v, a, 48); memcpy(B.v, b, 48); memcpy(Cc.v, c, 48); memcpy(D.v, d, 48);
  Fq381 r = fq381_mul_sub2(A, B, Cc, D);
  memcpy(out_mont, r.v, 48);
}

// generator of the radix-2 domain of size 2^logn over BLS12-381 Fr (csrc/ring.cuh), canonical limbs out
#include "../../ark_ec_vrfs_b200/csrc/ring.cuh"
extern "C" void hostemu_ntt_domain_gen(int logn, int inverse, uint32_t* out_canonical) {
  Fr255 w = ntt_domain_gen(logn, inverse != 0);
  from_mont<BlsFr>(out_canonical, w);
}
extern "C" void hostemu_ntt_inv_n(int logn, uint32_t* out_canonical) {
  from_mont<BlsFr>(out_canonical, fr_pow_u32(ntt_const(2), (uint32_t)logn));
}

// BLS12-381 G1 wire format (csrc/ring.cuh) on the host
extern "C" void hostemu_g1_compress(const uint8_t* pt96, uint8_t* out48) { g1_compress_one(out48, pt96); }
extern "C" int hostemu_g1_decompress(const uint8_t* in48, int check_subgroup, uint8_t* out96) { return g1_decompress_one(out96, in48, check_subgroup != 0); }

// BLS12-381 pairing (csrc/pairing.cuh) on the host: tower operations and the whole product check, ABI bytes in and out
#include "../../ark_ec_vrfs_b200/csrc/pairing_coop.cuh"
static Fq12 f12_from_bytes(const uint8_t* b) {
  Fq381 c[12]; uint32_t raw[12];
  for (int k = 0; k < 12; k++) { load_le<12>(raw, b + 48 * k); c[k] = to_mont<BlsFq>(raw); }
  return Fq12{Fq6{Fq2{c[0], c[1]}, Fq2{c[2], c[3]}, Fq2{c[4], c[5]}}, Fq6{Fq2{c[6], c[7]}, Fq2{c[8], c[9]}, Fq2{c[10], c[11]}}};
}
// op: 0 mul, 1 sqr, 2 inv, 3 frobenius q, 4 frobenius q^2, 5 cyclotomic squaring, 6 a * (b.c0.c0 + b.c0.c1 v + b.c1.c1 v w) sparse, 7 conj, 8 a^x, 9 final exponentiation
extern "C" void hostemu_f12_op(int op, const uint8_t* a576, const uint8_t* b576, uint8_t* out576) {
  Fq12 a = f12_from_bytes(a576), b = f12_from_bytes(b576), r = f12_one();
  switch (op) {
    case 0: r = f12_mul(a, b); break;
    case 1: r = f12_sqr(a); break;
    case 2: r = f12_inv(a); break;
    case 3: r = f12_frob<1>(a); break;
    case 4: r = f12_frob<2>(a); break;
    case 5: r = f12_cyclotomic_sqr(a); break;
    case 6: r = f12_mul_by_014(a, b.c0.c0, b.c0.c1, b.c1.c1); break;
    case 7: r = f12_conj(a); break;
    case 8: r = f12_exp_by_x(a); break;
    case 9: r = final_exponentiation(a); break;
  }
  f12_store(out576, r);
}
// the warp-cooperative form (csrc/pairing_coop.cuh): the interpreter of the generated lane programs, all lanes played in turn
extern "C" int hostemu_pairing_product_lanes(const uint8_t* g1, const uint8_t* g2, unsigned negate, uint8_t* out_gt576) {
  Fq12 e = f12_one();
  int verdict = pairing_product_check_lanes_host(g1, g2, negate, &e);
  if (out_gt576) f12_store(out_gt576, e);
  return verdict;
}
extern "C" int hostemu_pairing_product(int n, const uint8_t* g1, const uint8_t* g2, unsigned negate, uint8_t* out_gt576) {
  Fq12 e = f12_one();
  int verdict = pairing_product_check_bytes(n, g1, g2, negate, &e);
  if (out_gt576) f12_store(out_gt576, e);
  return verdict;
}
