// Unsaturated-limb arithmetic for the Bandersnatch base field (BLS12-381 Fr) on the hot path.
//
// Why: on B200 a carry-chained IMAD.WIDE.U32.X issues at 0.41 of the rate of a plain IMAD.WIDE
// (profiles/r1a_imad_microbench.json: 7.7 vs 18.5 T MAC32/s).  A saturated 8x32-bit Montgomery product is
// 136 carry-chained multiplies; here a field element is 9 SIGNED limbs of 29 bits, products accumulate into
// 64-bit columns with plain IMAD.WIDE (81 + 72 per product, p_0 = 1 and -1/p = -1 mod 2^29 cost nothing),
// and all carry handling is shifts/adds on the ALU pipe, which runs beside the multiplier.
//
// Representation: x = sum v[i] * 2^(29 i), "Montgomery" with R = 2^261, values are LAZY: any integer
// congruent to x*R mod p inside the stated bounds, possibly negative.
//   normalised : v[0..7] in [0, 2^29), v[8] a small signed remainder.
//   products   : f29_mul(a, b) needs max|a_i| * max|b_j| <= 2^59 (e.g. 2^30 x 2^29) and |a|*|b| <= 70 p^2;
//                it returns a normalised value in (-|ab|/R, |ab|/R + p).
//   a + b, a - b are limb-wise (no carries); f29_norm carries without changing the value.
// The bounds of every sequence in te29.cuh are listed there; tests/host_emul runs them on the host against
// big-integer arithmetic with random and extreme (all-ones / negated) limb patterns.
#pragma once
#include "gen/field_consts.cuh"

namespace vrfs {

struct alignas(4) F29 { int32_t v[9]; };
static constexpr uint32_t F29_MASK = (1u << 29) - 1u;

template <uint32_t (*Fn)(int)> HD_INLINE F29 f29c() { F29 r; for (int i = 0; i < 9; i++) r.v[i] = (int32_t)Fn(i); return r; }
HD_INLINE F29 f29_zero() { F29 r; for (int i = 0; i < 9; i++) r.v[i] = 0; return r; }
HD_INLINE F29 f29_one() { return f29c<F29Consts::ONE>(); }
HD_INLINE F29 f29_add(const F29& a, const F29& b) { F29 r; for (int i = 0; i < 9; i++) r.v[i] = a.v[i] + b.v[i]; return r; }
HD_INLINE F29 f29_sub(const F29& a, const F29& b) { F29 r; for (int i = 0; i < 9; i++) r.v[i] = a.v[i] - b.v[i]; return r; }
HD_INLINE F29 f29_neg(const F29& a) { F29 r; for (int i = 0; i < 9; i++) r.v[i] = -a.v[i]; return r; }
HD_INLINE F29 f29_dbl(const F29& a) { F29 r; for (int i = 0; i < 9; i++) r.v[i] = a.v[i] * 2; return r; }
HD_INLINE F29 f29_select(bool c, const F29& a, const F29& b) { F29 r; for (int i = 0; i < 9; i++) r.v[i] = c ? a.v[i] : b.v[i]; return r; }
HD_INLINE F29 f29_cneg(const F29& a, bool c) { F29 r; for (int i = 0; i < 9; i++) r.v[i] = c ? -a.v[i] : a.v[i]; return r; }
// carry propagation: same value, limbs 0..7 in [0, 2^29)
HD_INLINE F29 f29_norm(const F29& a) {
  F29 r;
  int32_t c = 0;
  for (int i = 0; i < 8; i++) { int32_t t = a.v[i] + c; r.v[i] = (int32_t)((uint32_t)t & F29_MASK); c = t >> 29; }
  r.v[8] = a.v[8] + c;
  return r;
}

// norm(5a + b) for NORMALISED a, b (limbs 0..7 non-negative): 5*a_i + b_i + carry < 6 * 2^29 + 8 fits an unsigned word,
// whereas 5*a_i alone already overflows int32
HD_INLINE F29 f29_norm_5a_plus_b(const F29& a, const F29& b) {
  F29 r;
  uint32_t c = 0;
  for (int i = 0; i < 8; i++) { uint32_t t = 5u * (uint32_t)a.v[i] + (uint32_t)b.v[i] + c; r.v[i] = (int32_t)(t & F29_MASK); c = t >> 29; }
  r.v[8] = 5 * a.v[8] + b.v[8] + (int32_t)c;
  return r;
}

// shared tail: Montgomery reduction of the 17 product columns t[0..16] (t[17] = 0 on entry) and normalisation
HD_INLINE F29 f29_reduce(int64_t* t) {
  int32_t P[9];
  for (int i = 0; i < 9; i++) P[i] = (int32_t)F29Consts::P(i);
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const uint32_t m = (0u - (uint32_t)t[i]) & F29_MASK;          // m * p_0 = m, and -1/p = -1 (mod 2^29)
    const int64_t c = (t[i] + (int64_t)m) >> 29;                  // exact: t[i] + m = 0 (mod 2^29)
#pragma unroll
    for (int j = 1; j < 9; j++) t[i + j] += (int64_t)(int32_t)m * P[j];
    t[i + 1] += c;
  }
  F29 r;
#pragma unroll
  for (int k = 9; k < 17; k++) {
    r.v[k - 9] = (int32_t)((uint32_t)t[k] & F29_MASK);
    t[k + 1] += t[k] >> 29;
  }
  r.v[8] = (int32_t)t[17];
  return r;
}
HD_NOINLINE F29 f29_mul(F29 a, F29 b) {
  int64_t t[18];
#pragma unroll
  for (int k = 0; k < 18; k++) t[k] = 0;
#pragma unroll
  for (int i = 0; i < 9; i++)
#pragma unroll
    for (int j = 0; j < 9; j++) t[i + j] += (int64_t)a.v[i] * b.v[j];
  return f29_reduce(t);
}
// a normalised (|a_i| <= 2^29): cross terms use 2*a_i
HD_NOINLINE F29 f29_sqr(F29 a) {
  int64_t t[18];
  int32_t a2[9];
#pragma unroll
  for (int k = 0; k < 18; k++) t[k] = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) a2[i] = a.v[i] * 2;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    t[2 * i] += (int64_t)a.v[i] * a.v[i];
#pragma unroll
    for (int j = i + 1; j < 9; j++) t[i + j] += (int64_t)a2[i] * a.v[j];
  }
  return f29_reduce(t);
}

// ---- conversions with the saturated representation Fp<BlsFr> (Montgomery, R = 2^256, canonical) ----------
HD_INLINE F29 f29_from_fp(const Fp<BlsFr>& x) {
  F29 y;   // split the canonical 256-bit integer into 29-bit limbs, then rescale by 2^266 / 2^261
  for (int i = 0; i < 9; i++) {
    int bit = 29 * i, w = bit >> 5, sh = bit & 31;
    uint64_t v = x.v[w];
    if (w + 1 < 8) v |= (uint64_t)x.v[w + 1] << 32;
    y.v[i] = (int32_t)((uint32_t)(v >> sh) & F29_MASK);
  }
  return f29_mul(y, f29c<F29Consts::K_IN>());
}
// lazy value -> canonical representative in [0, p), as 9 non-negative 29-bit limbs
HD_INLINE void f29_canonical_limbs(uint32_t* out, const F29& a_) {
  // a in (-4p, 4p) after the rescaling product; add 4p, then subtract p while >= p (at most 7 times; done branch-free)
  F29 a = f29_norm(f29_add(a_, f29c<F29Consts::P4>()));
  for (int rep = 0; rep < 7; rep++) {
    F29 d = f29_norm(f29_sub(a, f29c<F29Consts::P>()));
    bool neg = d.v[8] < 0;
    a = f29_select(neg, a, d);
  }
  for (int i = 0; i < 9; i++) out[i] = (uint32_t)a.v[i];
}
HD_INLINE Fp<BlsFr> f29_to_fp(const F29& x) {
  F29 y = f29_mul(x, f29c<F29Consts::K_OUT>());      // value * 2^256 (mod p), lazy in (-p, 2p)
  uint32_t l[9];
  f29_canonical_limbs(l, y);
  Fp<BlsFr> r;
  for (int w = 0; w < 8; w++) {
    int bit = 32 * w, i = bit / 29, sh = bit % 29;       // bits [32w, 32w+32) live in limbs i, i+1 (and i+2)
    uint64_t v = (uint64_t)l[i] >> sh;
    int have = 29 - sh;
    if (i + 1 < 9) v |= (uint64_t)l[i + 1] << have;
    if (i + 2 < 9 && have + 29 < 32) v |= (uint64_t)l[i + 2] << (have + 29);
    r.v[w] = (uint32_t)v;
  }
  return r;
}
HD_INLINE bool f29_is_zero(const F29& x) {
  uint32_t l[9];
  f29_canonical_limbs(l, x);
  uint32_t o = 0;
  for (int i = 0; i < 9; i++) o |= l[i];
  return o == 0;
}

}  // namespace vrfs
