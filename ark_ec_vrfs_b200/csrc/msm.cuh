// K12 of SURVEY.md 2.5: Pippenger bucket MSM over BLS12-381 G1 for the ring-VRF KZG commitment
// (`ring` -> ring-proof `index`/`commit` -> ark-ec `VariableBaseMSM::msm`, named at
// /root/reference/src/lib.rs:13-17): n_columns scalar columns over one base vector.
// The result is a group element; its canonical affine form does not depend on how the sum is grouped, so
// the window size, signed digits and bucket order used here need not match ark-ec's.
//
// Pipeline (all on the context's stream):
//   k_msm_prep_bases      canonical LE affine -> Montgomery affine (identity = all zero)
//   k_msm_prepare         (prepared mode, once per SRS) Q[w][i] = 2^(c w) P_i, so all windows share one bucket set
//   k_msm_histogram       signed radix-2^c digits of every scalar; per-(segment, bucket) counts
//   k_msm_scan            exclusive scan of the counts inside each segment; lists buckets with > MSM_BIG entries
//   k_msm_scatter         point indices (+ sign bit) grouped by bucket (counting sort)
//   k_msm_accumulate      one thread per bucket: complete additions of its points
//   k_msm_accumulate_big  one block per oversized bucket (top window / skewed scalar columns such as ring selectors)
//   k_msm_rc / k_msm_wbits segment value sum_j j*B_j: row / column sums of the bucket matrix, then the two short weighted sums bit by bit
//   k_msm_final2          one warp per column: Horner over the windows (stateless mode only); partial or affine out
// Stateless mode (vrfs_msm_g1_bls12_381): segment = (column, window) over the GLV halves of the scalars (2n bases [P | -phi(P)],
// 128-bit scalars: half the windows, a Horner chain of 128 doublings).  Prepared mode (vrfs_msm_g1_prepare + _prepared, the analogue of
// RingContext holding the SRS): segment = column, no Horner chain; short SRS (n <= 2^14) skip the whole bucket pipeline for a table of
// the 128 multiples of every 2^(8w) P_i ("TABLE mode" below: k_msm_table_build / _sum / _reduce).
#pragma once
#include "lincomb.cuh"
#include "gen/ntt_consts.cuh"

namespace vrfs {

typedef SWPoint<G1Curve> G1Pt;
typedef Fp<BlsFq> Fq381;
struct G1Aff { Fq381 x, y; };   // Montgomery form; x = y = 0 encodes the identity ((0,0) is not on the curve)

struct MsmPlan {
  uint32_t n, ncol;
  int c, windows, nb;            // window bits, digit windows per scalar, buckets per segment (2^(c-1))
  int prepared;                  // 1: bases pre-multiplied by 2^(c*w) -> ONE bucket segment per column, no Horner
  int seg_windows;               // segments per column: `windows` (stateless) or 1 (prepared)
  int tpb;                       // threads cooperating on one bucket in k_msm_accumulate (power of two <= 32)
  uint32_t big;                  // buckets with more entries than this go to k_msm_accumulate_big
  int warp_agg;                  // histogram / scatter: one atomic per group of lanes that hit the same bucket (set by callers that know their columns repeat values)
  int rc_h;                      // segment reduction: columns H of the R x H bucket matrix summed by k_msm_rc (0: nb <= 256, k_msm_wbits takes the buckets directly)
};
VRFS_HD inline MsmPlan msm_plan(uint32_t n, uint32_t ncol, int prepared, int c_override = 0, int tpb_override = 0, int scalar_bits = 255) {
  MsmPlan p; p.n = n; p.ncol = ncol; p.prepared = prepared;
  int lg = 0; while ((1u << (lg + 1)) <= n) lg++;
  // prepared mode: all windows of a column share one bucket set, so a short top window (255 mod c bits) piles n entries onto
  // 2^(255 mod c) buckets; the window sizes below keep that remainder at >= 5 bits (c = 12 and 14 left 3 bits: oversized
  // buckets cost 0.2-0.6 ms at 2^11..2^14)
  // tools/msm_sweep.py with scalars uniform below r (profiles/r1r_msm_sweep.json): c = 16 spends no window on the carry of bit 254
  if (prepared) p.c = lg <= 9 ? 8 : lg <= 12 ? 10 : lg <= 16 ? 13 : 16;
  // stateless mode (n counts the 2 x bases of the GLV halves, 128-bit scalars): tools/msm_sweep_stateless.py.  The top window must
  // not be a few bits wide (c = 9, 12 leave 2 / 8 bits: its buckets collect 8 x the average load); c = 8 spends it on the carry alone
  else p.c = lg <= 13 ? 8 : lg <= 14 ? 10 : lg <= 16 ? 11 : 13;
  if (c_override >= 7 && c_override <= 18) p.c = c_override;      // caller's window hint (vrfs_msm_g1_prepare_ex; tools/msm_sweep.py)
  p.windows = (scalar_bits + p.c) / p.c;  // scalar bits (255; 128 for the GLV halves of the stateless mode) + the top carry of the signed recoding
  p.nb = 1 << (p.c - 1);
  p.seg_windows = prepared ? 1 : p.windows;
  // small domains are latency-bound on chains of dependent additions: spread every stage over more threads
  // (>= 2 entries per thread in a bucket until the grid holds ~64 K threads: the in-bucket tree is cooperative and cheap)
  uint64_t avg = ((uint64_t)n * (prepared ? p.windows : 1)) / p.nb;      // expected entries per bucket for uniform digits
  const uint64_t total_buckets = (uint64_t)ncol * (prepared ? 1 : p.windows) * p.nb;
  p.tpb = 1; while (p.tpb < 32 && (avg / p.tpb > 24 || (total_buckets * p.tpb < 65536 && avg / p.tpb >= 2))) p.tpb *= 2;
  if (tpb_override >= 1 && tpb_override <= 32 && !(tpb_override & (tpb_override - 1))) p.tpb = tpb_override;
  p.big = (uint32_t)(3 * avg + 64);     // far above any natural load (Poisson tail); a padded ring's repeated point (N/4..N/2 entries per window) must land here
  p.warp_agg = 0;
  p.rc_h = 0;
  if (p.nb > 256) { int h = 0; while ((1 << (2 * h)) < p.nb) h++; p.rc_h = 1 << h; }   // H = 2^ceil(log2(nb)/2) <= 512
  return p;
}
#define MSM_BIG_THREADS 128
#define MSM_SLICE 256u           // an oversized bucket is cut into slices of about this many entries, one block each
#define MSM_MAX_SLICES 256u
// capacity of the slice list: buckets above p.big = 3 avg + 64 entries number < buckets / 3; every slice has >= MSM_SLICE entries but the last
VRFS_HD inline size_t msm_big_capacity(size_t nbuckets, size_t total_entries) { return nbuckets / 2 + total_entries / MSM_SLICE + 64; }

HD_INLINE void g1_load_aff(G1Pt& P, const G1Aff* a, bool negate) {
  uint4* d = reinterpret_cast<uint4*>(&P);
  const uint4* s = reinterpret_cast<const uint4*>(a);
  for (int i = 0; i < 6; i++) d[i] = s[i];
  bool inf = P.X.is_zero() & P.Y.is_zero();
  P.Y = cneg(P.Y, negate);
  P.Z = select(inf, Fq381::zero(), Fq381::one());
  P.Y = select(inf, Fq381::one(), P.Y);
}
HD_INLINE void g1_load_proj(G1Pt& P, const G1Pt* a, bool negate) {
  copy_words16(&P, a);
  P.Y = cneg(P.Y, negate);
}
// r = 2p, complete, a = 0 (RCB16 Algorithm 9: 6M + 2S + 1 mul_b3); checked against affine arithmetic in this container
HD_NOINLINE void g1_dbl(G1Pt* r, const G1Pt* p) {
  Fq381 t0 = sqr(p->Y), Z3 = dbl(dbl(dbl(t0))), t1 = p->Y * p->Z, t2 = G1Curve::mul_b3(sqr(p->Z));
  Fq381 X3 = t2 * Z3, Y3 = t0 + t2;
  Z3 = t1 * Z3;
  t0 = t0 - (dbl(t2) + t2);
  Y3 = t0 * Y3 + X3;
  X3 = dbl(t0 * (p->X * p->Y));
  r->X = X3; r->Y = Y3; r->Z = Z3;
}
// ---- bucket accumulation in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ; ZZ = 0 is the identity) with AFFINE table records:
// madd-2008-s costs 8M + 2S and ~7 field additions, against 12M and ~20 additions for the complete homogeneous formula.
// The formula is not complete, so the three exceptional cases are tested explicitly (they only occur for an empty
// accumulator, repeated bases and cancelling pairs - a warp almost never diverges on them).
struct G1Xyzz { Fq381 X, Y, ZZ, ZZZ; };
// a b - c d with ONE Montgomery reduction: the two 24-limb products are added as a b + c (q - d) < 2 q^2 < q 2^384 / 5 and reduced
// together (2 x 144 + 156 multiply-accumulates instead of 2 x 300).  VRFS_MSM_FUSED_RED=0 keeps the two separate products.
#ifndef VRFS_MSM_FUSED_RED
#define VRFS_MSM_FUSED_RED 1
#endif
HD_NOINLINE Fq381 fq381_mul_sub2(const Fq381& a, const Fq381& b, const Fq381& c, const Fq381& d) {
#if VRFS_MSM_FUSED_RED
  typedef MontChains<12> C;
  uint32_t t[24], t2[24], nd[12], pm[12];
  for (int i = 0; i < 12; i++) pm[i] = BlsFq::mod(i);
  C::sub(nd, pm, d.v);                                  // q - d in [1, q]: the reduction takes any value below q 2^384
  C::mul_wide(t, a.v, b.v);
  C::mul_wide(t2, c.v, nd);
  const uint32_t cy = C::add(t, t, t2);                 // 24-limb sum: low halves, then high halves with the carry
  uint32_t one[12];
  for (int i = 0; i < 12; i++) one[i] = 0;
  one[0] = cy;
  C::add(t + 12, t + 12, t2 + 12);
  C::add(t + 12, t + 12, one);
  Fq381 r;
  mont_reduce_wide<BlsFq>(r.v, t);
  return r;
#else
  return a * b - c * d;
#endif
}
HD_INLINE void xyzz_set_identity(G1Xyzz& a) { a.X = Fq381::zero(); a.Y = Fq381::zero(); a.ZZ = Fq381::zero(); a.ZZZ = Fq381::zero(); }
// (called, not inlined: with the accumulator's address taken it lives in local memory - 928 B of stack in k_msm_accumulate - but
//  inlining it and/or raising the register cap to 168 was measured at 3.07 / 3.07 / 3.14 / 3.16 ms for the 2^17 x 3 accumulate:
//  the kernel is bound by the multiplier pipe, not by that L1-resident traffic)
HD_NOINLINE void xyzz_madd(G1Xyzz* acc, const Fq381* x2, const Fq381* y2) {
  if (acc->ZZ.is_zero()) { acc->X = *x2; acc->Y = *y2; acc->ZZ = Fq381::one(); acc->ZZZ = Fq381::one(); return; }
  Fq381 U2 = *x2 * acc->ZZ, S2 = *y2 * acc->ZZZ;
  Fq381 P = U2 - acc->X, R = S2 - acc->Y;
  if (P.is_zero()) {
    if (R.is_zero()) {                       // same point: affine doubling (mdbl-2008-s-1)
      Fq381 U = dbl(*y2), V = sqr(U), W = U * V, S = *x2 * V, xx = sqr(*x2), M = dbl(xx) + xx;
      Fq381 X3 = sqr(M) - dbl(S);
      acc->Y = M * (S - X3) - W * *y2; acc->X = X3; acc->ZZ = V; acc->ZZZ = W;
    } else {
      xyzz_set_identity(*acc);               // P + (-P)
    }
    return;
  }
  Fq381 PP = sqr(P), PPP = P * PP, Q = acc->X * PP;
  Fq381 X3 = sqr(R) - PPP - dbl(Q);
  acc->Y = fq381_mul_sub2(R, Q - X3, acc->Y, PPP);
  acc->X = X3;
  acc->ZZ = acc->ZZ * PP;
  acc->ZZZ = acc->ZZZ * PPP;
}
// -> homogeneous projective (X ZZZ : Y ZZ : ZZ ZZZ); identity -> (0 : 1 : 0)
HD_INLINE void xyzz_to_proj(G1Pt& r, const G1Xyzz& a) {
  bool inf = a.ZZ.is_zero();
  r.X = a.X * a.ZZZ;
  r.Y = select(inf, Fq381::one(), a.Y * a.ZZ);
  r.Z = a.ZZ * a.ZZZ;
}
// affine record (identity = all zero) -> (x, +-y); false for the identity
HD_INLINE bool g1_load_aff_xy(Fq381& x, Fq381& y, const G1Aff* a, bool negate) {
  G1Aff t;
  copy_words16(&t, a);
  x = t.x; y = cneg(t.y, negate);
  return !(t.x.is_zero() & t.y.is_zero());
}

// signed radix-2^c digit w of the canonical scalar k (8 limbs): digits in [-2^(c-1), 2^(c-1)]
HD_INLINE int msm_digit(const uint32_t* k, int w, int c, int& carry) {
  int bit = w * c;
  uint32_t limb = bit >> 5, sh = bit & 31;
  uint64_t v = limb < 8 ? k[limb] : 0u;
  if (limb + 1 < 8) v |= (uint64_t)k[limb + 1] << 32;
  int d = (int)((v >> sh) & ((1u << c) - 1u)) + carry;
  carry = d > (1 << (c - 1));
  if (carry) d -= (1 << c);
  return d;
}
HD_INLINE void msm_load_scalar(uint32_t* k, const uint8_t* p) {
  uint32_t raw[8];
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w; raw[4] = b.x; raw[5] = b.y; raw[6] = b.z; raw[7] = b.w;
  if (is_canonical<BlsFr>(raw)) { for (int i = 0; i < 8; i++) k[i] = raw[i]; }      // the common case: nothing to reduce
  else from_mont<BlsFr>(k, to_mont<BlsFr>(raw));  // reduce mod r like the reference's scalar decode
}

// ---- GLV for the stateless mode: phi(x, y) = (beta x, y) is multiplication by -z^2 on G1 (z the curve parameter; the constant of
// g1_in_subgroup), and r = z^4 - z^2 + 1, so every scalar below r is  k = k1 + q z^2  with k1 < z^2 < 2^128 and q < 2^128:
//     k P = k1 P + q (-phi(P)) = k1 (x, y) + q (beta x, -y).
// The stateless MSM (the literal `VariableBaseMSM::msm` signature) runs over the 2n bases [P_i | -phi(P_i)] with these 128-bit
// halves: as many bucket additions as before, but half the windows - half the bucket sets to reduce and a Horner chain of 128
// instead of 255 dependent doublings (0.47 of the call's ~1.5 ms at N = 2^11).  The division by z^2 is a multiplication by
// floor(2^256 / z^2) and at most one correction (checked on 10^5 random scalars when the constants were derived; the host tests
// check k = k1 + q z^2 and the bounds on this very code).
HD_INLINE void g1_glv_split(uint32_t* k1 /*4*/, uint32_t* q /*4*/, const uint32_t* k /*8, canonical < r*/) {
  constexpr uint32_t Z2[4] = {0x00000000u, 0x00000001u, 0x0001a402u, 0xac45a401u};                    // z^2
  constexpr uint32_t M[5] = {0xf6cfee2eu, 0x63f6e522u, 0xe01faaddu, 0x7c6becf1u, 0x00000001u};        // floor(2^256 / z^2)
  uint32_t prod[13];
  for (int i = 0; i < 13; i++) prod[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 5; j++) { const uint64_t t = (uint64_t)k[i] * M[j] + prod[i + j] + c; prod[i + j] = (uint32_t)t; c = t >> 32; }
    prod[i + 5] = (uint32_t)c;
  }
  uint32_t qe[4] = {prod[8], prod[9], prod[10], prod[11]};      // floor(k M / 2^256) in {q - 1, q}; q < 2^128
  uint32_t qz[8];
  for (int i = 0; i < 8; i++) qz[i] = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 4; j++) { const uint64_t t = (uint64_t)qe[i] * Z2[j] + qz[i + j] + c; qz[i + j] = (uint32_t)t; c = t >> 32; }
    qz[i + 4] = (uint32_t)c;
  }
  uint32_t rem[5];                                              // k - qe z^2 < 2 z^2 < 2^129
  { uint64_t b = 0; for (int i = 0; i < 5; i++) { const uint64_t t = (uint64_t)k[i] - qz[i] - b; rem[i] = (uint32_t)t; b = (t >> 32) & 1u; } }
  uint32_t sub[5];
  uint64_t b = 0;
  for (int i = 0; i < 5; i++) { const uint64_t t = (uint64_t)rem[i] - (i < 4 ? Z2[i] : 0u) - b; sub[i] = (uint32_t)t; b = (t >> 32) & 1u; }
  const bool fix = b == 0;                                      // rem >= z^2: one more z^2 fits
  uint64_t c = fix ? 1u : 0u;
  for (int i = 0; i < 4; i++) { k1[i] = fix ? sub[i] : rem[i]; const uint64_t t = (uint64_t)qe[i] + c; q[i] = (uint32_t)t; c = t >> 32; }
}

// 1/a in BLS12-381 Fq by the binary extended Euclid (at most 2*381 shift/subtract steps on 12-limb integers) instead of a
// Fermat power (~480 dependent 12-limb products): the final projective -> affine step of an MSM runs on ONE thread per
// column, so its latency is the whole call's tail (0.43 ms -> 0.16 ms).  Montgomery in, Montgomery out; 0 -> 0.
HD_INLINE void fq381_shr1(uint32_t* a) { for (int i = 0; i < 11; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31); a[11] >>= 1; }
HD_INLINE void fq381_halve_mod(uint32_t* x, const uint32_t* p) {            // x / 2 mod p, x < p
  uint32_t t[12];
  if (x[0] & 1u) { MontChains<12>::add(t, x, p); for (int i = 0; i < 12; i++) x[i] = t[i]; }   // < 2^382: no carry out
  fq381_shr1(x);
}
HD_NOINLINE Fq381 fq381_inv(const Fq381& a) {
  typedef MontChains<12> C;
  if (a.is_zero()) return a;
  uint32_t u[12], v[12], x1[12], x2[12], p[12], t[12];
  for (int i = 0; i < 12; i++) { u[i] = a.v[i]; p[i] = BlsFq::mod(i); v[i] = p[i]; x1[i] = i == 0; x2[i] = 0; }
  auto is_one = [](const uint32_t* z) { uint32_t o = z[0] ^ 1u; for (int i = 1; i < 12; i++) o |= z[i]; return o == 0; };
  for (int guard = 0; guard < 2 * 384 && !is_one(u) && !is_one(v); guard++) {
    while (!(u[0] & 1u)) { fq381_shr1(u); fq381_halve_mod(x1, p); }
    while (!(v[0] & 1u)) { fq381_shr1(v); fq381_halve_mod(x2, p); }
    if (C::sub(t, u, v) == 0) {                       // u >= v
      for (int i = 0; i < 12; i++) u[i] = t[i];
      if (C::sub(t, x1, x2)) C::add(t, t, p);
      for (int i = 0; i < 12; i++) x1[i] = t[i];
    } else {
      C::sub(t, v, u);
      for (int i = 0; i < 12; i++) v[i] = t[i];
      if (C::sub(t, x2, x1)) C::add(t, t, p);
      for (int i = 0; i < 12; i++) x2[i] = t[i];
    }
  }
  Fq381 x;                                            // plain inverse of the integer a*R: (aR)^-1 = a^-1 R^-1
  const uint32_t* r = is_one(u) ? x1 : x2;
  for (int i = 0; i < 12; i++) x.v[i] = r[i];
  return x * Fq381::r3();                             // a^-1 R^-1 * R^3 / R = a^-1 R
}

// 1/a in BLS12-381 Fq, second version: binary GCD on word-sized approximations (Pornin, "Optimized Binary GCD for Modular
// Inversion", 2020), 32-bit flavour.  Each of the at most 26 outer rounds (the loop leaves as soon as a = 0, typically after 18-20) reads the top 32 and the low 30 bits of (a, b) into two
// 62-bit words, runs 30 branch-free binary-GCD steps on them while recording the 2x2 update matrix (entries <= 2^30), and
// applies the matrix once to the 12-limb a, b (exact division by 2^30) and to u, v (division by 2^30 mod p with one
// Montgomery-style correction word).  ~1/4 of the instructions of fq381_inv and no data-dependent branch, so the 32 lanes of a
// warp can invert 32 different values in lock-step.  Montgomery in, Montgomery out; 0 -> 0.  If the loop does not end in
// (a, b) = (0, 1) - never observed; the round count follows the paper's 2*len - 1 bound - the result is recomputed by fq381_inv.
HD_INLINE void bingcd_lin2(uint32_t* out13, const uint32_t* a, int32_t f, const uint32_t* b, int32_t g) {   // a*f + b*g, two's complement
  long long cy = 0;
  for (int i = 0; i < 12; i++) {
    long long t = (long long)((unsigned long long)a[i] * (uint32_t)f) - (f < 0 ? (long long)((unsigned long long)a[i] << 32) : 0ll);
    t += (long long)((unsigned long long)b[i] * (uint32_t)g) - (g < 0 ? (long long)((unsigned long long)b[i] << 32) : 0ll);
    t += cy;
    out13[i] = (uint32_t)t; cy = t >> 32;
  }
  out13[12] = (uint32_t)cy;
}
HD_INLINE bool fq381_inv_bingcd(Fq381& out, const Fq381& x) {        // false: the loop did not end in (0, 1) (x = 0, or never)
  uint32_t a[12], b[12], u[12], v[12], pm[12];
  for (int i = 0; i < 12; i++) { a[i] = x.v[i]; pm[i] = BlsFq::mod(i); b[i] = pm[i]; u[i] = i == 0; v[i] = 0; }
  const uint32_t ninv30 = BlsFq::NINV & 0x3fffffffu;             // -p^-1 mod 2^30
#pragma unroll 1
  for (int round = 0; round < 26; round++) {
    // bit length n of a | b, window position sp = max(n - 32, 30)
    uint32_t topw = 0, anz = 0; int topi = 0;
    for (int i = 0; i < 12; i++) { const uint32_t w = a[i] | b[i]; if (w) { topw = w; topi = i; } anz |= a[i]; }
    if (anz == 0) break;                                          // a = 0: b = gcd and v is final (u, v carry no pending power of two)
#ifdef __CUDA_ARCH__
    const int lz = __clz((int)topw);
#else
    const int lz = topw ? __builtin_clz(topw) : 32;
#endif
    int sp = 32 * topi + 32 - lz - 32; if (sp < 30) sp = 30;
    const int q = sp >> 5, r = sp & 31;
    uint32_t alo = 0, ahi = 0, blo = 0, bhi = 0;
    for (int i = 0; i < 12; i++) { if (i == q) { alo = a[i]; blo = b[i]; } if (i == q + 1) { ahi = a[i]; bhi = b[i]; } }
    const uint32_t atop = r ? (alo >> r) | (ahi << (32 - r)) : alo, btop = r ? (blo >> r) | (bhi << (32 - r)) : blo;
    unsigned long long xa = (a[0] & 0x3fffffffu) | ((unsigned long long)atop << 30), xb = (b[0] & 0x3fffffffu) | ((unsigned long long)btop << 30);
    // the matrix rows travel as F = f + g * 2^32 in two's complement: one 64-bit subtraction / doubling updates both entries
    long long F0 = 1, F1 = 1ll << 32;
    for (int j = 0; j < 30; j++) {
      const bool odd = xa & 1u, sw = odd & (xa < xb);
      const unsigned long long ta = sw ? xb : xa, tb = sw ? xa : xb;
      const long long tF0 = sw ? F1 : F0, tF1 = sw ? F0 : F1;
      xa = (ta - (odd ? tb : 0ull)) >> 1; xb = tb;
      F0 = tF0 - (odd ? tF1 : 0ll);
      F1 = (long long)((unsigned long long)tF1 << 1);
    }
    int32_t f0 = (int32_t)(uint32_t)F0, f1 = (int32_t)(uint32_t)F1;            // |f|, |g| <= 2^30
    int32_t g0 = (int32_t)((F0 - (long long)f0) >> 32), g1 = (int32_t)((F1 - (long long)f1) >> 32);
    uint32_t na[13], nb[13];
    bingcd_lin2(na, a, f0, b, g0);
    bingcd_lin2(nb, a, f1, b, g1);
    const bool nega = na[12] >> 31, negb = nb[12] >> 31;
    {                                                             // a = |na| >> 30, b = |nb| >> 30
      uint32_t ca = nega, cb = negb;
      const uint32_t ma = nega ? 0xffffffffu : 0u, mb = negb ? 0xffffffffu : 0u;
      for (int i = 0; i < 13; i++) {
        unsigned long long t = (unsigned long long)(na[i] ^ ma) + ca; na[i] = (uint32_t)t; ca = (uint32_t)(t >> 32);
        t = (unsigned long long)(nb[i] ^ mb) + cb; nb[i] = (uint32_t)t; cb = (uint32_t)(t >> 32);
      }
      for (int i = 0; i < 12; i++) { a[i] = (na[i] >> 30) | (na[i + 1] << 2); b[i] = (nb[i] >> 30) | (nb[i + 1] << 2); }
    }
    if (nega) { f0 = -f0; g0 = -g0; }
    if (negb) { f1 = -f1; g1 = -g1; }
    // (u, v) <- (u f0 + v g0, u f1 + v g1) / 2^30 mod p
    for (int which = 0; which < 2; which++) {
      const int32_t f = which ? f1 : f0, g = which ? g1 : g0;
      const uint32_t t0 = u[0] * (uint32_t)f + v[0] * (uint32_t)g;
      const uint32_t qq = (t0 * ninv30) & 0x3fffffffu;
      uint32_t t[13];
      long long cy = 0;
      for (int i = 0; i < 12; i++) {
        long long w = (long long)((unsigned long long)u[i] * (uint32_t)f) - (f < 0 ? (long long)((unsigned long long)u[i] << 32) : 0ll);
        w += (long long)((unsigned long long)v[i] * (uint32_t)g) - (g < 0 ? (long long)((unsigned long long)v[i] << 32) : 0ll);
        w += cy;
        // + p[i] * qq, added as an unsigned quantity: split so that the signed sum cannot overflow
        const unsigned long long pq = (unsigned long long)pm[i] * qq;
        w += (long long)(pq & 0xffffffffu);
        t[i] = (uint32_t)w; cy = (w >> 32) + (long long)(pq >> 32);
      }
      t[12] = (uint32_t)cy;
      uint32_t r12[12];
      for (int i = 0; i < 12; i++) r12[i] = (t[i] >> 30) | (t[i + 1] << 2);      // in (-p, 2p), two's complement in 384 bits
      const bool neg = t[12] >> 31;
      uint32_t add[12], tmp[12];
      for (int i = 0; i < 12; i++) add[i] = neg ? pm[i] : 0u;
      MontChains<12>::add(r12, r12, add);                                          // now in [0, 2p)
      const uint32_t borrow = MontChains<12>::sub(tmp, r12, pm);
      for (int i = 0; i < 12; i++) r12[i] = borrow ? r12[i] : tmp[i];
      if (which == 0) { for (int i = 0; i < 12; i++) na[i] = r12[i]; } else { for (int i = 0; i < 12; i++) nb[i] = r12[i]; }
    }
    for (int i = 0; i < 12; i++) { u[i] = na[i]; v[i] = nb[i]; }
  }
  uint32_t bad = b[0] ^ 1u;
  for (int i = 0; i < 12; i++) bad |= a[i] | (i ? b[i] : 0u);
  Fq381 r;                                                        // v = (xR)^-1 -> x^-1 R = v * R^3 / R
  for (int i = 0; i < 12; i++) r.v[i] = v[i];
  out = r * Fq381::r3();
  return bad == 0;
}
HD_NOINLINE Fq381 fq381_inv_fast(const Fq381& x) {
  Fq381 r;
  if (fq381_inv_bingcd(r, x)) return r;
  return fq381_inv(x);                                            // x = 0 (-> 0)
}

#ifdef __CUDACC__
// =================================================================================================
// Segment reduction (second generation): everything after the bucket sums is a chain of
// DEPENDENT point additions on very few points, and one thread needs ~11 us per complete addition (12 products of
// 12 limbs on a multiplier pipe that retires one 64-bit product per ~6 cycles per warp).  Two ideas:
//  * lane-cooperative arithmetic: a complete addition is 6 independent products, a few additions, 6 more independent
//    products.  A group of 8 lanes holds the operands replicated; lane g computes product g of each level and the six
//    results are exchanged with shuffles: ~3 us per addition / doubling instead of ~11 (g1_coop_add, g1_coop_dbl).
//  * sum_j j*B_j as (1) row and column sums of the R x H bucket matrix (j - 1 = hi*H + lo;
//    sum j*B_j = H * sum_hi hi*Row_hi + sum_lo (lo+1)*Col_lo: two adds per bucket, depth log2 H, k_msm_rc) and
//    (2) the two short weighted sums by the bits of their weights (k_msm_wbits, round 2: one block per bit gathers the points
//    whose weight has that bit, doubles the sum bit-many times; round 1 ran a halving recursion on an 8-block cluster, 2.5 x
//    slower at these sizes): no multiplication by chunk offsets and no Horner chain.
// (First generation - one thread per chunk of buckets with a double-and-add by the chunk offset, a block tree, a one-thread final -
//  took 1.11 ms at 2^17 x 3 against 0.41 ms now; it is in the history before session 4.)
// =================================================================================================
// 1/x in F_q by FOUR lanes: the word-approximation GCD of fq381_inv_bingcd with its four 12-limb updates of a round
// (a, b <- exact quotients; u, v <- quotients mod q) spread over the lanes of a group - lane r of a group holds ONE of a, b, u, v -
// and run through one uniform code path (a lane-dependent code path would be serialised by the warp): every lane forms
// X f + Y g + qq q over 13 limbs (qq = 0 for a and b), shifts by 30 bits, then one masked fix-up (|.| for a, b; + q when negative
// for u, v), one conditional - q, and one conditional negation mod q of u / v by the sign its a / b partner found.  The 30-step
// inner loop on the 62-bit approximations is computed by all lanes alike.  ~750 instructions per round and lane instead of
// ~1000 on one thread (measured: k_msm_final2 0.065 -> 0.055 ms; the 30-step inner loop is most of a round either way).
// Every lane of the warp must call it; x is the same in the four lanes of a group; all lanes of a group return 1/x (0 -> 0;
// the fallback for a loop that did not end in (0, 1) - never observed - is the one-thread routine).
__device__ __noinline__ Fq381 fq381_inv_coop4(const Fq381& x, bool* fell_back = nullptr) {
  const unsigned lane = threadIdx.x & 31u, role = lane & 3u, base = lane & ~3u;
  const bool is_ab = role < 2u, second = (role & 1u) != 0;     // roles: 0 a, 1 b, 2 u, 3 v
  uint32_t V[12], pm[12];
  for (int i = 0; i < 12; i++) {
    pm[i] = BlsFq::mod(i);
    V[i] = role == 0 ? x.v[i] : role == 1 ? pm[i] : role == 2 ? (i == 0 ? 1u : 0u) : 0u;
  }
  const uint32_t ninv30 = BlsFq::NINV & 0x3fffffffu;
#pragma unroll 1
  for (int round = 0; round < 26; round++) {
    uint32_t O[12];                                            // the partner's value: a <-> b, u <-> v
    for (int i = 0; i < 12; i++) O[i] = __shfl_xor_sync(0xffffffffu, V[i], 1);
    const uint32_t* X = second ? O : V;                        // (X, Y) = (a, b) or (u, v)
    const uint32_t* Y = second ? V : O;
    uint32_t topw = 0, anz = 0; int topi = 0;
    for (int i = 0; i < 12; i++) { const uint32_t w = X[i] | Y[i]; if (w) { topw = w; topi = i; } anz |= X[i]; }
    anz = __shfl_sync(0xffffffffu, anz, base);                 // a of this group
    if (!__any_sync(0xffffffffu, anz != 0)) break;             // every group of the warp is done (a finished group idles correctly: a stays 0, b and v stay)
    const int lz = __clz((int)topw);
    int sp = 32 * topi + 32 - lz - 32; if (sp < 30) sp = 30;
    const int q = sp >> 5, r = sp & 31;
    uint32_t alo = 0, ahi = 0, blo = 0, bhi = 0;
    for (int i = 0; i < 12; i++) { if (i == q) { alo = X[i]; blo = Y[i]; } if (i == q + 1) { ahi = X[i]; bhi = Y[i]; } }
    const uint32_t atop = r ? (alo >> r) | (ahi << (32 - r)) : alo, btop = r ? (blo >> r) | (bhi << (32 - r)) : blo;
    unsigned long long xa = (X[0] & 0x3fffffffu) | ((unsigned long long)atop << 30), xb = (Y[0] & 0x3fffffffu) | ((unsigned long long)btop << 30);
    xa = __shfl_sync(0xffffffffu, xa, base); xb = __shfl_sync(0xffffffffu, xb, base);        // the (a, b) approximations, for all four lanes
    long long F0 = 1, F1 = 1ll << 32;
    for (int j = 0; j < 30; j++) {
      const bool odd = xa & 1u, sw = odd & (xa < xb);
      const unsigned long long ta = sw ? xb : xa, tb = sw ? xa : xb;
      const long long tF0 = sw ? F1 : F0, tF1 = sw ? F0 : F1;
      xa = (ta - (odd ? tb : 0ull)) >> 1; xb = tb;
      F0 = tF0 - (odd ? tF1 : 0ll);
      F1 = (long long)((unsigned long long)tF1 << 1);
    }
    const long long Fm = second ? F1 : F0;                     // this lane's row of the matrix
    const int32_t f = (int32_t)(uint32_t)Fm;
    const int32_t g = (int32_t)((Fm - (long long)f) >> 32);
    // t = X f + Y g + qq q (13 limbs, two's complement); qq makes the low 30 bits vanish for u, v (they do by themselves for a, b)
    const uint32_t t0 = X[0] * (uint32_t)f + Y[0] * (uint32_t)g;
    const uint32_t qq = is_ab ? 0u : ((t0 * ninv30) & 0x3fffffffu);
    uint32_t t[13];
    long long cy = 0;
    for (int i = 0; i < 12; i++) {
      long long w = (long long)((unsigned long long)X[i] * (uint32_t)f) - (f < 0 ? (long long)((unsigned long long)X[i] << 32) : 0ll);
      w += (long long)((unsigned long long)Y[i] * (uint32_t)g) - (g < 0 ? (long long)((unsigned long long)Y[i] << 32) : 0ll);
      w += cy;
      const unsigned long long pq = (unsigned long long)pm[i] * qq;
      w += (long long)(pq & 0xffffffffu);
      t[i] = (uint32_t)w; cy = (w >> 32) + (long long)(pq >> 32);
    }
    t[12] = (uint32_t)cy;
    uint32_t r12[12];
    for (int i = 0; i < 12; i++) r12[i] = (t[i] >> 30) | (t[i + 1] << 2);      // a, b: exact, |.| < 2^383; u, v: in (-q, 2q)
    const bool negv = (t[12] >> 31) != 0;
    {  // a, b: |r|;  u, v: r + q when negative
      const uint32_t xm = (is_ab && negv) ? 0xffffffffu : 0u;
      uint32_t add[12];
      for (int i = 0; i < 12; i++) add[i] = (!is_ab && negv) ? pm[i] : 0u;
      if (is_ab && negv) add[0] = 1u;
      for (int i = 0; i < 12; i++) r12[i] ^= xm;
      MontChains<12>::add(r12, r12, add);
    }
    {  // u, v: - q when >= q
      uint32_t tmp[12];
      const uint32_t borrow = MontChains<12>::sub(tmp, r12, pm);
      const bool take = !is_ab && borrow == 0;
      for (int i = 0; i < 12; i++) r12[i] = take ? tmp[i] : r12[i];
    }
    {  // the sign a / b lost goes to u / v:  u <- -u mod q  (lane r + 2 follows lane r)
      const bool flip = __shfl_xor_sync(0xffffffffu, (int)negv, 2) != 0 && !is_ab;
      uint32_t tmp[12], nz = 0;
      MontChains<12>::sub(tmp, pm, r12);
      for (int i = 0; i < 12; i++) nz |= r12[i];
      for (int i = 0; i < 12; i++) r12[i] = (flip && nz) ? tmp[i] : r12[i];
    }
    for (int i = 0; i < 12; i++) V[i] = r12[i];
  }
  // (a, b) must have ended in (0, 1); v = (x R)^-1
  uint32_t bad = 0;
  if (role == 0) for (int i = 0; i < 12; i++) bad |= V[i];
  if (role == 1) { bad = V[0] ^ 1u; for (int i = 1; i < 12; i++) bad |= V[i]; }
  bad |= __shfl_xor_sync(0xffffffffu, bad, 1);
  bad = __shfl_sync(0xffffffffu, bad, base);
  Fq381 res;
  for (int i = 0; i < 12; i++) res.v[i] = __shfl_sync(0xffffffffu, V[i], base + 3);
  res = res * Fq381::r3();
  if (fell_back) *fell_back = bad != 0;
  if (__any_sync(0xffffffffu, bad != 0)) { if (bad) res = fq381_inv(x); }   // x = 0 (-> 0), or the never-observed unfinished loop
  return res;
}
__device__ __forceinline__ Fq381 fq_shfl(const Fq381& a, unsigned src_lane) {
  Fq381 r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src_lane);
  return r;
}
// r = p + q (complete, the formula of sw_add with a = 0) computed by the 8 lanes of a group; p, q and r are replicated in every
// lane of the group.  All 32 lanes of the warp must call it together.
__device__ __noinline__ void g1_coop_add(G1Pt* r, const G1Pt* p, const G1Pt* q) {
  const unsigned lane = threadIdx.x & 31u, g = lane & 7u, base = lane & ~7u;
  Fq381 a, b;
  {
    // level 1, lane g: 0 X1*X2, 1 Y1*Y2, 2 Z1*Z2, 3 (X1+Y1)(X2+Y2), 4 (X1+Z1)(X2+Z2), 5 (Y1+Z1)(Y2+Z2)   (6, 7 repeat 0)
    const bool fx = (g == 0) | (g == 3) | (g == 4) | (g >= 6), fy = (g == 1) | (g == 5);
    const bool sy = (g == 3), sz = (g == 4) | (g == 5);
    a = select(fx, p->X, select(fy, p->Y, p->Z)) + select(sy, p->Y, select(sz, p->Z, Fq381::zero()));
    b = select(fx, q->X, select(fy, q->Y, q->Z)) + select(sy, q->Y, select(sz, q->Z, Fq381::zero()));
  }
  Fq381 t = a * b;
  Fq381 t0 = fq_shfl(t, base), t1 = fq_shfl(t, base + 1), t2 = fq_shfl(t, base + 2);
  Fq381 t3 = fq_shfl(t, base + 3) - (t0 + t1);            // X1Y2 + X2Y1
  Fq381 t4 = fq_shfl(t, base + 4) - (t0 + t2);            // X1Z2 + X2Z1
  Fq381 t5 = fq_shfl(t, base + 5) - (t1 + t2);            // Y1Z2 + Y2Z1
  Fq381 Z3 = G1Curve::mul_b3(t2), X3 = t1 - Z3;
  Z3 = t1 + Z3;
  t1 = dbl(t0) + t0;
  t4 = G1Curve::mul_b3(t4);
  // level 2, lane g: 0 X3*Z3, 1 t1*t4, 2 t3*X3, 3 t5*t4, 4 t5*Z3, 5 t3*t1
  a = select(g == 0, X3, select(g == 1, t1, select((g == 2) | (g == 5), t3, t5)));
  b = select((g == 0) | (g == 4), Z3, select((g == 1) | (g == 3), t4, select(g == 2, X3, t1)));
  t = a * b;
  r->Y = fq_shfl(t, base) + fq_shfl(t, base + 1);
  r->X = fq_shfl(t, base + 2) - fq_shfl(t, base + 3);
  r->Z = fq_shfl(t, base + 4) + fq_shfl(t, base + 5);
}
// r = 2p (the formula of g1_dbl), same conventions
__device__ __noinline__ void g1_coop_dbl(G1Pt* r, const G1Pt* p) {
  const unsigned lane = threadIdx.x & 31u, g = lane & 7u, base = lane & ~7u;
  // level 1, lane g: 0 Y*Y, 1 Y*Z, 2 Z*Z, 3 X*Y
  Fq381 a = select(g == 2, p->Z, select(g == 3, p->X, p->Y));
  Fq381 b = select((g == 1) | (g == 2), p->Z, p->Y);
  Fq381 t = a * b;
  Fq381 t0 = fq_shfl(t, base), t1 = fq_shfl(t, base + 1), t2 = G1Curve::mul_b3(fq_shfl(t, base + 2)), xy = fq_shfl(t, base + 3);
  Fq381 Z3 = dbl(dbl(dbl(t0))), Y3 = t0 + t2;
  t0 = t0 - (dbl(t2) + t2);
  // level 2, lane g: 0 t2*Z3, 1 t1*Z3, 2 t0*Y3, 3 t0*xy
  a = select(g == 0, t2, select(g == 1, t1, t0));
  b = select((g == 0) | (g == 1), Z3, select(g == 2, Y3, xy));
  t = a * b;
  r->Y = fq_shfl(t, base + 2) + fq_shfl(t, base);
  r->X = dbl(fq_shfl(t, base + 3));
  r->Z = fq_shfl(t, base + 1);
}

// sh[0] = sum of sh[0 .. width) (width a power of two <= blockDim.x = 128, entries already stored and the block synchronised):
// one thread per addition while more than 16 remain, then groups of 8 lanes per addition
__device__ __forceinline__ void block_tree_sum(G1Pt* sh, int width) {
  const unsigned t = threadIdx.x;
  for (int s = width >> 1; s > 16; s >>= 1) {
    if ((int)t < s) { G1Pt x, y; copy_words16(&x, &sh[t]); copy_words16(&y, &sh[t + s]); sw_add<G1Curve>(&x, &x, &y); copy_words16(&sh[t], &x); }
    __syncthreads();
  }
  const unsigned grp = t >> 3, g = t & 7u;                      // 16 groups of 8 lanes
  for (int s = min(width >> 1, 16); s > 0; s >>= 1) {
    if ((int)(grp & ~3u) < s) {                                 // warp-uniform: the warp holds a live group
      G1Pt x, y, z;
      copy_words16(&x, &sh[(int)grp < s ? grp : (unsigned)s]);            // idle groups of a live warp read what nobody writes
      copy_words16(&y, &sh[(int)grp < s ? grp + s : (unsigned)s]);
      g1_coop_add(&z, &x, &y);
      __syncwarp();
      if ((int)grp < s && g == 0) copy_words16(&sh[grp], &z);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128) k_msm_prep_bases(uint32_t n, const uint8_t* bases, G1Aff* out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* q = reinterpret_cast<const uint4*>(bases + (size_t)96 * i);
  uint32_t rx[12], ry[12];
  for (int j = 0; j < 3; j++) { uint4 a = q[j]; rx[4 * j] = a.x; rx[4 * j + 1] = a.y; rx[4 * j + 2] = a.z; rx[4 * j + 3] = a.w; }
  for (int j = 0; j < 3; j++) { uint4 a = q[3 + j]; ry[4 * j] = a.x; ry[4 * j + 1] = a.y; ry[4 * j + 2] = a.z; ry[4 * j + 3] = a.w; }
  G1Aff o;
  o.x = to_mont<BlsFq>(rx); o.y = to_mont<BlsFq>(ry);
  out[i] = o;
}
// stateless mode: bases2 = [P_i | -phi(P_i)], scalars2[col] = [k1_i | q_i] as 32-byte values (see g1_glv_split)
__global__ void __launch_bounds__(128) k_msm_glv_split(uint32_t n, uint32_t ncol, const G1Aff* aff, const uint8_t* scalars, G1Aff* bases2, uint8_t* scalars2) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * ncol) return;
  const uint32_t col = t / n, i = t % n;
  uint32_t k[8], k1[4], q[4];
  msm_load_scalar(k, scalars + (size_t)32 * t);
  g1_glv_split(k1, q, k);
  uint4* o1 = reinterpret_cast<uint4*>(scalars2 + (size_t)32 * ((size_t)col * 2 * n + i));
  uint4* o2 = reinterpret_cast<uint4*>(scalars2 + (size_t)32 * ((size_t)col * 2 * n + n + i));
  o1[0] = make_uint4(k1[0], k1[1], k1[2], k1[3]); o1[1] = make_uint4(0, 0, 0, 0);
  o2[0] = make_uint4(q[0], q[1], q[2], q[3]); o2[1] = make_uint4(0, 0, 0, 0);
  if (col == 0) {
    G1Aff a; copy_words16(&a, &aff[i]);
    copy_words16(&bases2[i], &a);
    Fq381 beta;
    for (int j = 0; j < 12; j++) beta.v[j] = G1WireConsts::beta(j);
    a.x = a.x * beta; a.y = neg(a.y);                          // the identity (0, 0) stays (0, 0)
    copy_words16(&bases2[n + i], &a);
  }
}
// prepared bases (the RingContext analogue: the SRS is fixed): Q[w*n + i] = 2^(c*w) * P_i, AFFINE (96 B; identity = zeros).
// One thread per base: the doubling chain leaves projective points, whose Z's are inverted together (Montgomery's trick, one
// binary-Euclid inversion per base).
#define MSM_MAX_WINDOWS 33
__global__ void __launch_bounds__(128) k_msm_prepare(uint32_t n, int c, int windows, const G1Aff* aff, G1Aff* Q) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Pt P;
  g1_load_aff(P, &aff[i], false);
  Fq381 zs[MSM_MAX_WINDOWS], pre[MSM_MAX_WINDOWS];
  unsigned long long infmask = 0;                                // windows whose point is the identity (Z = 0)
  for (int w = 0; w < windows; w++) {
    if (w) for (int k = 0; k < c; k++) g1_dbl(&P, &P);
    G1Aff t; t.x = P.X; t.y = P.Y;                               // unnormalised for now
    copy_words16(&Q[(size_t)w * n + i], &t);
    const bool inf = P.Z.is_zero();
    if (inf) infmask |= 1ull << w;
    zs[w] = select(inf, Fq381::one(), P.Z);                      // an identity must not poison the product chain
    pre[w] = w ? pre[w - 1] * zs[w] : zs[w];
  }
  Fq381 acc = fq381_inv_fast(pre[windows - 1]);
  for (int w = windows - 1; w >= 0; w--) {
    Fq381 zi = w ? acc * pre[w - 1] : acc;
    if (w) acc = acc * zs[w];
    G1Aff t; copy_words16(&t, &Q[(size_t)w * n + i]);
    t.x = t.x * zi; t.y = t.y * zi;
    if ((infmask >> w) & 1ull) { t.x = Fq381::zero(); t.y = Fq381::zero(); }
    copy_words16(&Q[(size_t)w * n + i], &t);
  }
}
// counts[seg*nb + (|d|-1)]++ ; seg = col*seg_windows + (prepared ? 0 : w)
// p.warp_agg: lanes of a warp that hit the same bucket in the same window (a ring's repeated padding point, its 0/1 selector) are
// found with match.any and served by ONE atomic.  Measured at 2^17: padded Lagrange ring 3.66 -> 3.41 ms (histogram 0.18 -> 0.08,
// scatter 0.23 -> 0.08 ms), random columns 3.82 -> 3.90 ms - so only the ring entry points, which know their columns repeat, set it.
__global__ void __launch_bounds__(128) k_msm_histogram(MsmPlan p, const uint8_t* scalars, uint32_t* counts) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
  const bool valid = t < p.n * p.ncol;
  const uint32_t col = valid ? t / p.n : 0u;
  uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (valid) msm_load_scalar(k, scalars + (size_t)32 * t);
  int carry = 0;
  for (int w = 0; w < p.windows; w++) {
    const int d = msm_digit(k, w, p.c, carry);
    const size_t seg = (size_t)col * p.seg_windows + (p.prepared ? 0 : w);
    const size_t b = seg * p.nb + (d < 0 ? -d : d) - 1;
    const bool on = valid && d != 0;
    if (p.warp_agg) {
      const unsigned m = __match_any_sync(0xffffffffu, on ? (unsigned long long)b : (0xffffffff00000000ull | lane));
      if (on && lane == (unsigned)(__ffs(m) - 1)) atomicAdd(&counts[b], (uint32_t)__popc(m));
    } else if (on) atomicAdd(&counts[b], 1u);
  }
}
// one block per segment: exclusive scan of nb counts -> offsets (relative to the segment); also lists the big buckets
#define MSM_SCAN_THREADS 1024
__global__ void __launch_bounds__(MSM_SCAN_THREADS) k_msm_scan(MsmPlan p, const uint32_t* counts, uint32_t* offsets, uint32_t* big_list, uint32_t* big_count) {
  __shared__ uint32_t part[MSM_SCAN_THREADS];
  const uint32_t seg = blockIdx.x;
  const uint32_t* c = counts + (size_t)seg * p.nb;
  uint32_t* o = offsets + (size_t)seg * p.nb;
  const int per = (p.nb + MSM_SCAN_THREADS - 1) / MSM_SCAN_THREADS;
  const int lo = threadIdx.x * per, hi = min(lo + per, p.nb);
  uint32_t s = 0;
  for (int i = lo; i < hi; i++) s += c[i];
  // block-wide exclusive scan of the per-thread sums: warp scans by shuffle, then the 32 warp totals
  const uint32_t lane_ = threadIdx.x & 31u, warp_ = threadIdx.x >> 5;
  uint32_t incl = s;
  for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane_ >= d) incl += v; }
  if (lane_ == 31) part[warp_] = incl;
  __syncthreads();
  if (warp_ == 0) {
    uint32_t w = part[lane_], wi = w;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d); if ((int)lane_ >= d) wi += v; }
    part[lane_] = wi - w;
  }
  __syncthreads();
  uint32_t run = part[warp_] + incl - s;
  for (int i = lo; i < hi; i++) {
    o[i] = run; run += c[i];
    if (c[i] > p.big) {                                       // one work item per slice of ~MSM_SLICE entries
      // every slice ends in a block-wide tree of complete additions (~5 mixed additions' worth per thread), so long buckets get
      // longer slices: 256 entries up to 4096, 1024 above (4096-entry slices measured worse: too few blocks for a lone long bucket)
      const uint32_t slice = c[i] > 16u * MSM_SLICE ? 4u * MSM_SLICE : MSM_SLICE;
      const uint32_t nsl = min((uint32_t)MSM_MAX_SLICES, (c[i] + slice - 1) / slice);
      const uint32_t first = atomicAdd(big_count, nsl);
      for (uint32_t sl = 0; sl < nsl; sl++) { big_list[2 * (first + sl)] = seg * p.nb + i; big_list[2 * (first + sl) + 1] = sl | (nsl << 16); }
    }
  }
}
// list[seg*seg_len + offsets[bucket] + pos] = point index | sign << 31
__global__ void __launch_bounds__(128) k_msm_scatter(MsmPlan p, const uint8_t* scalars, const uint32_t* offsets, uint32_t* cursors, uint32_t* list) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
  const bool valid = t < p.n * p.ncol;
  const uint32_t col = valid ? t / p.n : 0u, i = valid ? t % p.n : 0u;
  uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (valid) msm_load_scalar(k, scalars + (size_t)32 * t);
  const size_t seg_len = (size_t)p.n * (p.prepared ? p.windows : 1);
  int carry = 0;
  for (int w = 0; w < p.windows; w++) {
    const int d = msm_digit(k, w, p.c, carry);
    const size_t seg = (size_t)col * p.seg_windows + (p.prepared ? 0 : w);
    const size_t b = seg * p.nb + (d < 0 ? -d : d) - 1;
    const bool on = valid && d != 0;
    uint32_t pos = 0;
    if (p.warp_agg) {
      const unsigned m = __match_any_sync(0xffffffffu, on ? (unsigned long long)b : (0xffffffff00000000ull | lane));
      const unsigned leader = (unsigned)(__ffs(m) - 1);
      uint32_t base = 0;
      if (on && lane == leader) base = atomicAdd(&cursors[b], (uint32_t)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      pos = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
    } else if (on) pos = atomicAdd(&cursors[b], 1u);
    if (!on) continue;
    const uint32_t idx = p.prepared ? (uint32_t)w * p.n + i : i;
    list[seg * seg_len + offsets[b] + pos] = idx | (d < 0 ? 0x80000000u : 0u);
  }
}
// every table record is affine: the caller's bases (stateless) or the prepared 2^(c w) P_i
__device__ __forceinline__ void msm_accumulate_entries(G1Pt& out, const void* bases, const uint32_t* l, uint32_t first, uint32_t cnt, uint32_t step) {
  const G1Aff* tab = reinterpret_cast<const G1Aff*>(bases);
  G1Xyzz acc; xyzz_set_identity(acc);
  for (uint32_t j = first; j < cnt; j += step) {
    const uint32_t e = l[j];
    Fq381 x, y;
    const bool finite = g1_load_aff_xy(x, y, tab + (e & 0x7fffffffu), (e >> 31) != 0);
    if (j + step < cnt) prefetch_l1(tab + (l[j + step] & 0x7fffffffu), (unsigned)sizeof(G1Aff));   // a random 96-byte record of a table far larger than L2
    if (finite) xyzz_madd(&acc, &x, &y);
  }
  xyzz_to_proj(out, acc);
}
// p.tpb threads per bucket (strided over its entries, shared-memory tree inside the group); buckets above p.big
// entries are left to k_msm_accumulate_big
#ifndef MSM_ACC_MINBLOCKS
#define MSM_ACC_MINBLOCKS 4   // 128 registers: 16 warps/SM instead of 8 (measured 4.59 -> 4.17 ms at 2^17 x 3)
#endif
template <bool PREP>
// (b_begin, b_end): the range of buckets this launch covers - large calls run the last buckets as a second launch with more threads per
// bucket on a side stream, so that the slots the main launch frees while it drains are filled with SHORT tasks (msm_dev)
__global__ void __launch_bounds__(128, MSM_ACC_MINBLOCKS) k_msm_accumulate(MsmPlan p, const void* bases, const uint32_t* counts, const uint32_t* offsets,
                                                         const uint32_t* list, G1Pt* buckets, size_t b_begin, size_t b_end) {
  __shared__ uint4 sh_raw[128 * sizeof(G1Pt) / 16];
  G1Pt* sh = reinterpret_cast<G1Pt*>(sh_raw);
  const size_t total = b_end;
  const uint32_t lane = threadIdx.x % p.tpb;
  const size_t b = b_begin + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / p.tpb;
  const bool live = b < total;
  uint32_t cnt = live ? counts[b] : 0u;
  if (cnt > p.big) cnt = 0;                       // handled by the big path (which also writes the bucket)
  G1Pt acc; sw_set_identity(acc);
  if (cnt) {
    const size_t seg = b / p.nb;
    const size_t seg_len = (size_t)p.n * (p.prepared ? p.windows : 1);
    const uint32_t* l = list + seg * seg_len + offsets[b];
    msm_accumulate_entries(acc, bases, l, lane, cnt, (uint32_t)p.tpb);
  }
  if (p.tpb > 1) {
    // tree over the tpb partial sums of every bucket of the block: the first level (64 additions per block) one thread per
    // addition, the later ones (32, 16, ... additions per block) by the block's 16 cooperative groups of 8 lanes
    copy_words16(&sh[threadIdx.x], &acc);
    __syncthreads();
    int stride = p.tpb >> 1;
    if ((int)lane < stride) { G1Pt y; copy_words16(&y, &sh[threadIdx.x + stride]); sw_add<G1Curve>(&acc, &acc, &y); copy_words16(&sh[threadIdx.x], &acc); }
    __syncthreads();
    const unsigned grp = threadIdx.x >> 3, g = threadIdx.x & 7u;
    for (stride >>= 1; stride > 0; stride >>= 1) {
      const int adds = (128 / p.tpb) * stride;                   // <= 32
      for (int a0 = 0; a0 < adds; a0 += 16) {
        if ((int)(a0 + (grp & ~3u)) < adds) {                    // warp-uniform: the warp holds a live group
          const int a = a0 + (int)grp;
          const bool on = a < adds;
          const int slot = on ? (a / stride) * p.tpb + (a % stride) : stride;      // idle groups read what nobody writes at this level
          G1Pt x, y, z;
          copy_words16(&x, &sh[slot]); copy_words16(&y, &sh[on ? slot + stride : slot]);
          g1_coop_add(&z, &x, &y);
          __syncwarp();
          if (on && g == 0) copy_words16(&sh[slot], &z);
        }
      }
      __syncthreads();
    }
    if (lane == 0) copy_words16(&acc, &sh[threadIdx.x]);
  }
  if (live && lane == 0 && counts[b] <= p.big) copy_words16(&buckets[b], &acc);
}
// one block per slice of an oversized bucket: strided partial sums + shared-memory tree -> bigpart[slice];
// k_msm_big_combine then adds the slices of each bucket.  (A ring's 0/1 selector column puts half the domain into ONE bucket:
// a single block needed 4.8 ms for it at N = 2^17.)
template <bool PREP>
__global__ void __launch_bounds__(MSM_BIG_THREADS) k_msm_accumulate_big(MsmPlan p, const void* bases, const uint32_t* counts, const uint32_t* offsets,
                                                                         const uint32_t* list, const uint32_t* big_list, const uint32_t* big_count, G1Pt* bigpart) {
  __shared__ uint4 sh_raw[MSM_BIG_THREADS * sizeof(G1Pt) / 16];
  G1Pt* sh = reinterpret_cast<G1Pt*>(sh_raw);
  const uint32_t nbig = *big_count;
  for (uint32_t bi = blockIdx.x; bi < nbig; bi += gridDim.x) {
    const size_t b = big_list[2 * bi];
    const uint32_t sl = big_list[2 * bi + 1] & 0xffffu, nsl = big_list[2 * bi + 1] >> 16;
    const size_t seg = b / p.nb;
    const size_t seg_len = (size_t)p.n * (p.prepared ? p.windows : 1);
    const uint32_t cnt = counts[b];
    const uint32_t lo = (uint32_t)(((uint64_t)cnt * sl) / nsl), hi = (uint32_t)(((uint64_t)cnt * (sl + 1)) / nsl);
    const uint32_t* l = list + seg * seg_len + offsets[b] + lo;
    G1Pt acc;
    msm_accumulate_entries(acc, bases, l, threadIdx.x, hi - lo, MSM_BIG_THREADS);
    copy_words16(&sh[threadIdx.x], &acc);
    __syncthreads();
    block_tree_sum(sh, MSM_BIG_THREADS);
    if (threadIdx.x == 0) copy_words16(&bigpart[bi], &sh[0]);
    __syncthreads();
  }
}
__global__ void __launch_bounds__(MSM_BIG_THREADS) k_msm_big_combine(const uint32_t* big_list, const uint32_t* big_count, const G1Pt* bigpart, G1Pt* buckets) {
  __shared__ uint4 sh_raw[MSM_BIG_THREADS * sizeof(G1Pt) / 16];
  G1Pt* sh = reinterpret_cast<G1Pt*>(sh_raw);
  const uint32_t nbig = *big_count;
  for (uint32_t bi = blockIdx.x; bi < nbig; bi += gridDim.x) {
    if ((big_list[2 * bi + 1] & 0xffffu) != 0) continue;        // block-uniform: only the first slice of a bucket leads
    const uint32_t nsl = big_list[2 * bi + 1] >> 16;
    G1Pt acc; sw_set_identity(acc);
    for (uint32_t j = threadIdx.x; j < nsl; j += MSM_BIG_THREADS) { G1Pt q; copy_words16(&q, &bigpart[bi + j]); sw_add<G1Curve>(&acc, &acc, &q); }
    copy_words16(&sh[threadIdx.x], &acc);
    __syncthreads();
    int width = MSM_BIG_THREADS; while (width / 2 >= (int)nsl) width /= 2;
    block_tree_sum(sh, width);
    if (threadIdx.x == 0) copy_words16(&buckets[big_list[2 * bi]], &sh[0]);
    __syncthreads();
  }
}
// row and column sums of the bucket matrix of every segment: out[seg][hi] = sum_lo B[hi*H + lo] (hi < R = nb / H),
// out[seg][R + lo] = sum_hi B[hi*H + lo].  One block per sum; the last five tree levels (<= 16 additions) are cooperative.
__global__ void __launch_bounds__(128) k_msm_rc(MsmPlan p, const G1Pt* buckets, G1Pt* out) {
  __shared__ uint4 sh_raw[128 * sizeof(G1Pt) / 16];
  G1Pt* sh = reinterpret_cast<G1Pt*>(sh_raw);
  const int H = p.rc_h, R = p.nb / H;
  const uint32_t seg = blockIdx.y, o = blockIdx.x, t = threadIdx.x;
  const G1Pt* B = buckets + (size_t)seg * p.nb;
  const bool row = (int)o < R;
  const int count = row ? H : R, stride = row ? 1 : H;
  const G1Pt* first = row ? B + (size_t)o * H : B + (o - R);
  G1Pt acc; sw_set_identity(acc);
  if ((int)t < count) copy_words16(&acc, first + (size_t)t * stride);
  for (int k = t + 128; k < count; k += 128) { G1Pt q; copy_words16(&q, first + (size_t)k * stride); sw_add<G1Curve>(&acc, &acc, &q); }
  copy_words16(&sh[t], &acc);
  __syncthreads();
  int width = 128; while (width / 2 >= count) width /= 2;       // live entries of sh (power of two >= min(count, 128))
  block_tree_sum(sh, width);
  if (t == 0) copy_words16(&out[(size_t)seg * (R + H) + o], &sh[0]);
}

// The weighted sums  W = sum_i (i + shift) * x_i, i < m <= 512  (shift = 1: bucket values / column sums; shift = 0: the row index
// of k_msm_rc, followed by `post` = log2 H doublings) by the BITS of the weights:
//   W = sum_b 2^b S_b,   S_b = sum of the x_i whose weight has bit b set.
// One block per (bit, part, segment): its 32 cooperating groups pick up the <= m/2 selected points (one coop addition per 32
// points), a 5-level tree in shared memory gives S_b, and the first warp doubles it b (+ post) times; k_msm_final2 adds the
// msm_wbits(p) results of every part.  No cluster barriers and no chain longer than log2(nb) doublings: 0.04 ms where the
// halving recursion W(x) = W(y) + sum_u x_{2u+1}, y_u = 2 (x_{2u} + x_{2u+1}) run by an 8-block cluster (round 1; in the history) took
// 0.10 ms (m = 16 / 32 at N = 2^11), and 0.08 instead of 0.17 ms at N = 2^17.
// (weights <= 512 < 2^10: at most ten results per weighted sum)
VRFS_HD inline int msm_wbits(const MsmPlan& p) {           // bits of the largest weight of a plan's weighted sums
  const int wmax = p.rc_h ? p.rc_h : p.nb;                 // columns / buckets carry weights 1 .. H (or nb); rows 0 .. R - 1 < H
  int b = 0; while ((1 << b) <= wmax) b++;
  return b;
}
// stateless mode: the parts of every (column, window) segment are added by their own warp before the Horner chain of k_msm_final2
__global__ void __launch_bounds__(32) k_msm_sum_parts(int nparts, const G1Pt* parts, G1Pt* out) {
  const G1Pt* W = parts + (size_t)blockIdx.x * nparts;
  const unsigned grp = threadIdx.x >> 3;
  G1Pt sum; sw_set_identity(sum);
  for (int j0 = 0; j0 < nparts; j0 += 4) {
    const int j = j0 + (int)grp;
    G1Pt q; sw_set_identity(q);
    if (j < nparts) copy_words16(&q, &W[j]);
    if (j0 == 0) sum = q; else g1_coop_add(&sum, &sum, &q);
  }
  for (int s = 2; s > 0; s >>= 1) {
    G1Pt other;
    other.X = fq_shfl(sum.X, (threadIdx.x + 8u * s) & 31u); other.Y = fq_shfl(sum.Y, (threadIdx.x + 8u * s) & 31u); other.Z = fq_shfl(sum.Z, (threadIdx.x + 8u * s) & 31u);
    g1_coop_add(&sum, &sum, &other);
  }
  if (threadIdx.x == 0) copy_words16(&out[blockIdx.x], &sum);
}
__global__ void __launch_bounds__(256) k_msm_wbits(MsmPlan p, const G1Pt* in, G1Pt* parts) {
  __shared__ uint4 sh_raw[32 * sizeof(G1Pt) / 16];
  G1Pt* sh = reinterpret_cast<G1Pt*>(sh_raw);
  const uint32_t b = blockIdx.x, part = blockIdx.y, nparts = gridDim.y, seg = blockIdx.z;
  const unsigned G = threadIdx.x >> 3, g = threadIdx.x & 7u;
  const G1Pt* x; int m, shift, post = 0;
  if (p.rc_h) {
    const int R = p.nb / p.rc_h;
    x = in + (size_t)seg * (R + p.rc_h) + (part ? R : 0);
    m = part ? p.rc_h : R; shift = part ? 1 : 0;
    if (!part) while ((1 << post) < p.rc_h) post++;
  } else { x = in + (size_t)seg * p.nb; m = p.nb; shift = 1; }
  G1Pt* out = parts + ((size_t)seg * nparts + part) * gridDim.x + b;           // gridDim.x = msm_wbits(p) bits per part
  // the k-th weight with bit b set is w_k = (k >> b) << (b + 1) | 1 << b | (k & (2^b - 1)); weights run over [shift, m + shift)
  const int top = m + shift;                               // exclusive
  int K = 0;
  {
    const int full = top >> (b + 1), rem = top & ((1 << (b + 1)) - 1);
    K = full * (1 << b) + (rem > (1 << b) ? rem - (1 << b) : 0);
  }
  G1Pt acc; sw_set_identity(acc);
  if (K == 0) { if (threadIdx.x == 0) copy_words16(out, &acc); return; }
  for (int r = 0; r * 32 < K; r++) {
    const int k = r * 32 + (int)G;
    G1Pt e; sw_set_identity(e);
    if (k < K) {
      const int w = ((k >> b) << (b + 1)) | (1 << b) | (k & ((1 << b) - 1));
      copy_words16(&e, x + (w - shift));
    }
    if (r == 0) acc = e;
    else if ((r * 32 + (int)(G & ~3u)) < K) g1_coop_add(&acc, &acc, &e);      // warp-uniform: some group of this warp has a point
  }
  if (g == 0) copy_words16(&sh[G], &acc);
  __syncthreads();
  int live = 1; while (live < K && live < 32) live <<= 1;
  for (int s = live >> 1; s > 0; s >>= 1) {
    if ((int)(G & ~3u) < s) {
      G1Pt a2, b2, z;
      copy_words16(&a2, &sh[(int)G < s ? G : (unsigned)s]);
      copy_words16(&b2, &sh[(int)G < s ? G + s : (unsigned)s]);
      g1_coop_add(&z, &a2, &b2);
      __syncwarp();
      if ((int)G < s && g == 0) copy_words16(&sh[G], &z);
    }
    __syncthreads();
  }
  if (G < 4) {                                             // the first warp; every group computes the same
    G1Pt z; copy_words16(&z, &sh[0]);
    for (int k = 0; k < (int)b + post; k++) g1_coop_dbl(&z, &z);
    __syncwarp();
    if (G == 0 && g == 0) copy_words16(out, &z);
  }
}

// =================================================================================================
// TABLE mode of a prepared SRS (round 2): for the short domains ring commitments actually have (ring size <= 2^11) the bucket
// pipeline is bound by its nine dependent launches and by the weighted bucket sums, not by the additions.  With the SRS fixed,
// memory buys them off: T[w][i][d-1] = d 2^(8w) P_i for d = 1 .. 128 (affine, 96 B; 393 KB per base - 0.8 GB for a 2^11-point SRS
// of the 180 GB), every scalar is 32 signed radix-256 digits (digit_w = byte_w(k + 0x80..80) - 128, no carry chain), and a
// commitment is the plain SUM of the n x 32 table entries its digits select: no sort, no buckets, no weights.
//   k_msm_table_build   (once per SRS) the 128 multiples of every Q[w][i] = 2^(8w) P_i, normalised 16 at a time (Montgomery's trick)
//   k_msm_table_sum     every thread adds the entries of its strided share of the (window, base) pairs of one column (XYZZ mixed
//                       additions, the next entry prefetched), block tree -> one partial per block
//   k_msm_table_reduce  one block per column adds the blocks' partials;  k_msm_final2 normalises / exchanges as before
// Ring-shaped columns (a repeated padding point, a 0/1 selector) cost exactly what random columns cost: there is no bucket to overfill.
// =================================================================================================
#define MSM_TABLE_C 8
#define MSM_TABLE_D (1 << (MSM_TABLE_C - 1))              // 128 multiples per (window, base)
#define MSM_TABLE_WINDOWS 32                               // signed radix-256 digits of a scalar < r < 2^255 (the bias addition cannot overflow: 0x73 + 0x80 < 0x100)
#define MSM_TABLE_CHUNK 16
__global__ void __launch_bounds__(128) k_msm_table_build(uint32_t n, const G1Aff* Q /*[32][n]*/, G1Aff* T /*[32][n][128]*/) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // = w * n + i
  if (idx >= (size_t)MSM_TABLE_WINDOWS * n) return;
  Fq381 bx, by;
  const bool finite = g1_load_aff_xy(bx, by, Q + idx, false);
  G1Aff* out = T + idx * MSM_TABLE_D;
  G1Xyzz cur; xyzz_set_identity(cur);
  G1Xyzz st[MSM_TABLE_CHUNK];
  Fq381 pre[MSM_TABLE_CHUNK];
#pragma unroll 1
  for (int d0 = 0; d0 < MSM_TABLE_D; d0 += MSM_TABLE_CHUNK) {
    unsigned infmask = 0;
    for (int j = 0; j < MSM_TABLE_CHUNK; j++) {
      if (finite) xyzz_madd(&cur, &bx, &by);
      st[j] = cur;
      const bool inf = cur.ZZ.is_zero();                 // a base of small order (or the identity itself): that multiple is the identity
      if (inf) infmask |= 1u << j;
      const Fq381 wj = select(inf, Fq381::one(), cur.ZZ * cur.ZZZ);
      pre[j] = j ? pre[j - 1] * wj : wj;
    }
    Fq381 acc = fq381_inv_fast(pre[MSM_TABLE_CHUNK - 1]);
    for (int j = MSM_TABLE_CHUNK - 1; j >= 0; j--) {
      const bool inf = (infmask >> j) & 1u;
      const Fq381 wj = select(inf, Fq381::one(), st[j].ZZ * st[j].ZZZ);
      const Fq381 wi = j ? acc * pre[j - 1] : acc;       // 1 / (ZZ ZZZ)
      if (j) acc = acc * wj;
      G1Aff a;
      a.x = st[j].X * (wi * st[j].ZZZ);                  // X / ZZ
      a.y = st[j].Y * (wi * st[j].ZZ);                   // Y / ZZZ
      if (inf) { a.x = Fq381::zero(); a.y = Fq381::zero(); }
      copy_words16(&out[d0 + j], &a);
    }
  }
}
// blockIdx.y = column; the blocks of a column stride together over its n * 32 (window, base) pairs
__global__ void __launch_bounds__(128, 4) k_msm_table_sum(uint32_t n, const G1Aff* T, const uint8_t* scalars, G1Pt* partials) {
  __shared__ uint4 sh_raw[128 * sizeof(G1Pt) / 16];
  G1Pt* sh = reinterpret_cast<G1Pt*>(sh_raw);
  const uint32_t col = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  const uint32_t M = n * MSM_TABLE_WINDOWS;
  const uint8_t* sc = scalars + (size_t)32 * n * col;
  G1Xyzz acc; xyzz_set_identity(acc);
  // term m = w * n + i; its table entry: the digit of scalar i in window w
  auto entry_of = [&](uint32_t m, bool& negate) -> const G1Aff* {
    const uint32_t w = m / n, i = m - w * n;
    uint32_t k[9];
    msm_load_scalar(k, sc + (size_t)32 * i);
    k[8] = 0;
    add_window_bias<9>(k, 0x80808080u, 8);
    const int d = digit8(k, (int)w);
    negate = d < 0;
    const int mag = d < 0 ? -d : d;
    return mag ? T + ((size_t)m * MSM_TABLE_D + (mag - 1)) : nullptr;
  };
  bool neg = false;
  const G1Aff* e = t < M ? entry_of(t, neg) : nullptr;
  if (e) prefetch_l1(e, (unsigned)sizeof(G1Aff));
  for (uint32_t m = t; m < M; m += stride) {
    bool nneg = false;
    const G1Aff* ne = m + stride < M ? entry_of(m + stride, nneg) : nullptr;
    if (ne) prefetch_l1(ne, (unsigned)sizeof(G1Aff));          // a random 96-byte record of a table far larger than L2
    if (e) {
      Fq381 x, y;
      if (g1_load_aff_xy(x, y, e, neg)) xyzz_madd(&acc, &x, &y);
    }
    e = ne; neg = nneg;
  }
  G1Pt r;
  xyzz_to_proj(r, acc);
  copy_words16(&sh[threadIdx.x], &r);
  __syncthreads();
  block_tree_sum(sh, 128);
  if (threadIdx.x == 0) copy_words16(&partials[(size_t)col * gridDim.x + blockIdx.x], &sh[0]);
}
// one block of 256 threads (32 cooperating groups) per column: sum of its nparts partials -> out[col]
__global__ void __launch_bounds__(256) k_msm_table_reduce(int nparts, const G1Pt* partials, G1Pt* out) {
  __shared__ uint4 sh_raw[32 * sizeof(G1Pt) / 16];
  G1Pt* sh = reinterpret_cast<G1Pt*>(sh_raw);
  const G1Pt* W = partials + (size_t)blockIdx.x * nparts;
  const unsigned G = threadIdx.x >> 3, g = threadIdx.x & 7u;
  G1Pt acc; sw_set_identity(acc);
  for (int r = 0; r * 32 < nparts; r++) {
    const int k = r * 32 + (int)G;
    G1Pt e; sw_set_identity(e);
    if (k < nparts) copy_words16(&e, W + k);
    if (r == 0) acc = e;
    else if ((r * 32 + (int)(G & ~3u)) < nparts) g1_coop_add(&acc, &acc, &e);
  }
  if (g == 0) copy_words16(&sh[G], &acc);
  __syncthreads();
  int live = 1; while (live < nparts && live < 32) live <<= 1;
  for (int s = live >> 1; s > 0; s >>= 1) {
    if ((int)(G & ~3u) < s) {
      G1Pt a2, b2, z;
      copy_words16(&a2, &sh[(int)G < s ? G : (unsigned)s]);
      copy_words16(&b2, &sh[(int)G < s ? G + s : (unsigned)s]);
      g1_coop_add(&z, &a2, &b2);
      __syncwarp();
      if ((int)G < s && g == 0) copy_words16(&sh[G], &z);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) copy_words16(&out[blockIdx.x], &sh[0]);
}

// ---- multi-GPU exchange of the per-rank partial sums ("MSM splits by point range, with only a tiny NCCL/NVLink
// reduction of partial bucket sums").  Every rank owns a MAILBOX in its device memory that the other ranks of the node have mapped
// (in-process: peer access; across processes: CUDA IPC).  The final kernel of a rank's MSM stores its projective partial straight
// into the peers' mailboxes over NVLink, raises a flag behind a system-scope fence, then - on the ranks that fold - waits for the
// flags of all ranks in ITS OWN memory, adds the partials and normalises: the collective is part of the kernel that produced the
// data, no NCCL call, no host hop (an all-gather of 144 B per column per rank).
// Slots are double-buffered by call parity: a rank can only be one collective ahead of the slowest one (it needs that rank's
// partial to finish), so parity e and e+2 never overlap.
#define VRFS_MAX_PEERS 16
#define VRFS_PEER_MAXCOL 32
struct PeerBox {                                           // layout of one rank's mailbox
  uint32_t slot[2][VRFS_MAX_PEERS][VRFS_PEER_MAXCOL][36];  // [parity][sender][column] X, Y, Z (Montgomery limbs, identical on every GPU)
  unsigned long long flag[2][VRFS_MAX_PEERS][VRFS_PEER_MAXCOL];   // epoch of the last partial stored there
};
struct PeerArgs {
  int rank, world, root;           // root < 0: every rank folds (all-gather); else only `root` does (gather)
  unsigned long long epoch;        // this collective's number (> 0, the same on every rank)
  unsigned long long timeout_ns;   // a rank that waits longer gives up (status word), so a lost peer cannot hang the GPU
  PeerBox* box[VRFS_MAX_PEERS];    // every rank's mailbox as mapped here (box[rank] = this GPU's own)
  unsigned int* timed_out;         // mapped pinned host word, set to 1 on a timeout
};
__device__ __forceinline__ unsigned long long peer_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned long long peer_ld_acquire(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void peer_st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// one warp per column: segment value = sum of its parts, Horner over the windows (stateless mode).
// out_mode 0: affine canonical bytes (96 B); 1: projective canonical bytes (144 B, the host-gathered multi-GPU form);
//          2: publish the partial to the peers' mailboxes and (on folding ranks) fold all ranks' partials -> affine bytes (96 B)
__global__ void __launch_bounds__(32) k_msm_final2(MsmPlan p, int nparts, const G1Pt* parts, uint8_t* out, int out_mode, PeerArgs peer) {
  const uint32_t col = blockIdx.x;
  const G1Pt* W = parts + (size_t)col * p.seg_windows * nparts;
  G1Pt acc; sw_set_identity(acc);
  const unsigned grp = threadIdx.x >> 3;                  // the warp's four groups of eight lanes add the parts of a window side by side
  for (int w = p.seg_windows - 1; w >= 0; w--) {
    if (w != p.seg_windows - 1) for (int k = 0; k < p.c; k++) g1_coop_dbl(&acc, &acc);
    if (nparts == 1) { G1Pt q; copy_words16(&q, &W[w]); g1_coop_add(&acc, &acc, &q); continue; }      // stateless: summed by k_msm_sum_parts
    G1Pt part_sum; sw_set_identity(part_sum);
    for (int j0 = 0; j0 < nparts; j0 += 4) {
      const int j = j0 + (int)grp;
      G1Pt q; sw_set_identity(q);
      if (j < nparts) copy_words16(&q, &W[(size_t)w * nparts + j]);
      if (j0 == 0) part_sum = q; else g1_coop_add(&part_sum, &part_sum, &q);
    }
    // fold the four groups' sums: lanes 8..31 hand theirs down (two shuffle-exchange levels)
    for (int s = 2; s > 0; s >>= 1) {
      G1Pt other;
      other.X = fq_shfl(part_sum.X, (threadIdx.x + 8u * s) & 31u); other.Y = fq_shfl(part_sum.Y, (threadIdx.x + 8u * s) & 31u); other.Z = fq_shfl(part_sum.Z, (threadIdx.x + 8u * s) & 31u);
      g1_coop_add(&part_sum, &part_sum, &other);
    }
    // every group now holds the same total (the additions commute; all four computed a rotation of the same sum)
    g1_coop_add(&acc, &acc, &part_sum);
  }
  if (out_mode == 2) {
    const unsigned lane = threadIdx.x;
    const int par = (int)(peer.epoch & 1ull);
    // publish: lane r < world (or lane 0 -> root) stores the 36 words + the flag into rank r's mailbox
    const bool sends = peer.root < 0 ? (int)lane < peer.world : lane == 0;
    if (sends) {
      const int dst = peer.root < 0 ? (int)lane : peer.root;
      uint32_t* s = peer.box[dst]->slot[par][peer.rank][col];
      uint4* d4 = reinterpret_cast<uint4*>(s);
      const uint4 *sx = reinterpret_cast<const uint4*>(acc.X.v), *sy = reinterpret_cast<const uint4*>(acc.Y.v), *sz = reinterpret_cast<const uint4*>(acc.Z.v);
      for (int i = 0; i < 3; i++) { d4[i] = sx[i]; d4[3 + i] = sy[i]; d4[6 + i] = sz[i]; }
      __threadfence_system();
      peer_st_release(&peer.box[dst]->flag[par][peer.rank][col], peer.epoch);
    }
    if (peer.root >= 0 && peer.root != peer.rank) return;
    // fold: wait for every rank's flag in this GPU's own mailbox, then add (sender order is fixed, so every rank adds in the same order)
    PeerBox* mine = peer.box[peer.rank];
    // The wait is ONE warp-uniform loop (lane r polls rank r's flag, the vote ends it for all lanes at once): a per-lane loop leaves
    // the warp split into the groups that left it at different times, and everything after it - the fold, the four-lane inversion -
    // then issues once per group (measured: + 0.1 ms per call on two GPUs).
    bool late = false;
    {
      const unsigned long long t0 = peer_globaltimer();
      unsigned spins = 0;
      for (;;) {
        const bool ready = (int)lane >= peer.world || peer_ld_acquire(&mine->flag[par][lane][col]) == peer.epoch;
        if (__all_sync(0xffffffffu, ready)) break;
        if ((++spins & 1023u) == 0 && peer_globaltimer() - t0 > peer.timeout_ns) { late = true; break; }
      }
    }
    late = __any_sync(0xffffffffu, late);
    __syncwarp();
    if (late) {                                            // give up loudly: zeros out, host sees the status word
      if (lane == 0) { *peer.timed_out = 1u; __threadfence_system(); for (int i = 0; i < 96; i++) out[(size_t)96 * col + i] = 0; }
      return;
    }
    G1Pt sum; sw_set_identity(sum);
    for (int r = 0; r < peer.world; r++) {
      G1Pt q;
      const uint4* s4 = reinterpret_cast<const uint4*>(mine->slot[par][r][col]);
      uint4 *qx = reinterpret_cast<uint4*>(q.X.v), *qy = reinterpret_cast<uint4*>(q.Y.v), *qz = reinterpret_cast<uint4*>(q.Z.v);
      for (int i = 0; i < 3; i++) { qx[i] = __ldcv(s4 + i); qy[i] = __ldcv(s4 + 3 + i); qz[i] = __ldcv(s4 + 6 + i); }   // written by another GPU: never from L1
      g1_coop_add(&sum, &sum, &q);
    }
    // every lane takes lane 0's view of the total: only the lanes that polled a flag have an acquire behind their loads of the
    // mailbox, and the four-lane inversion below needs the same Z in all lanes of a group
    acc.X = fq_shfl(sum.X, 0); acc.Y = fq_shfl(sum.Y, 0); acc.Z = fq_shfl(sum.Z, 0);
  }
  uint32_t raw[12];
  if (out_mode == 1) {
    if (threadIdx.x != 0) return;
    uint8_t* o = out + (size_t)144 * col;
    from_mont<BlsFq>(raw, acc.X); store_le<12>(o, raw);
    from_mont<BlsFq>(raw, acc.Y); store_le<12>(o + 48, raw);
    from_mont<BlsFq>(raw, acc.Z); store_le<12>(o + 96, raw);
  } else {
    uint8_t* o = out + (size_t)96 * col;
    const Fq381 zi = fq381_inv_coop4(acc.Z);         // the whole warp (acc is replicated in every lane); identity: Z = 0 -> zi = 0 -> zeros
    if (threadIdx.x == 0) { from_mont<BlsFq>(raw, acc.X * zi); store_le<12>(o, raw); }
    if (threadIdx.x == 1) { from_mont<BlsFq>(raw, acc.Y * zi); store_le<12>(o + 48, raw); }
  }
}
// fold partial sums of several ranks: partials[part][col] projective LE canonical -> affine out
__global__ void k_g1_sum_partials(int n_parts, int ncol, const uint8_t* partials, uint8_t* out) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  G1Pt acc; sw_set_identity(acc);
  for (int part = 0; part < n_parts; part++) {
    const uint8_t* s = partials + ((size_t)part * ncol + col) * 144;
    uint32_t raw[12];
    G1Pt q;
    load_le<12>(raw, s); q.X = to_mont<BlsFq>(raw);
    load_le<12>(raw, s + 48); q.Y = to_mont<BlsFq>(raw);
    load_le<12>(raw, s + 96); q.Z = to_mont<BlsFq>(raw);
    sw_add<G1Curve>(&acc, &acc, &q);
  }
  uint32_t raw[12];
  uint8_t* o = out + (size_t)96 * col;
  Fq381 zi = fq381_inv_fast(acc.Z);
  from_mont<BlsFq>(raw, acc.X * zi); store_le<12>(o, raw);
  from_mont<BlsFq>(raw, acc.Y * zi); store_le<12>(o + 48, raw);
}
#endif  // __CUDACC__

}  // namespace vrfs
