// Scalar-side helpers of K7/K8: byte <-> limb conversion, GLV splitting for
// Bandersnatch (k = k1 + k2*lambda, |k1|,|k2| < 2^127), signed fixed-window recoding.
// None of this changes results: k*P is the same group element whichever way the scalar is split,
// and the affine result is canonical (the reference uses bit-serial `mul_bigint`).
#pragma once
#include "gen/field_consts.cuh"

namespace vrfs {

template <int N> HD_INLINE void load_le(uint32_t* r, const uint8_t* b) {
  for (int i = 0; i < N; i++) r[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
}
template <int N> HD_INLINE void store_le(uint8_t* b, const uint32_t* r) {
  for (int i = 0; i < N; i++) { b[4 * i] = (uint8_t)r[i]; b[4 * i + 1] = (uint8_t)(r[i] >> 8); b[4 * i + 2] = (uint8_t)(r[i] >> 16); b[4 * i + 3] = (uint8_t)(r[i] >> 24); }
}
// big-endian bytes (4N of them) -> little-endian limbs
template <int N> HD_INLINE void load_be(uint32_t* r, const uint8_t* b) {
  for (int i = 0; i < N; i++) { const uint8_t* p = b + 4 * (N - 1 - i); r[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
}
template <int N> HD_INLINE void store_be(uint8_t* b, const uint32_t* r) {
  for (int i = 0; i < N; i++) { uint8_t* p = b + 4 * (N - 1 - i); p[0] = (uint8_t)(r[i] >> 24); p[1] = (uint8_t)(r[i] >> 16); p[2] = (uint8_t)(r[i] >> 8); p[3] = (uint8_t)r[i]; }
}

// r[0..NA+NB-1] = a * b (schoolbook on 64-bit temporaries; used a handful of times per item)
template <int NA, int NB> HD_INLINE void mul_wide(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  for (int i = 0; i < NA + NB; i++) r[i] = 0;
  for (int i = 0; i < NA; i++) {
    uint64_t c = 0;
    for (int j = 0; j < NB; j++) { uint64_t t = (uint64_t)a[i] * b[j] + r[i + j] + c; r[i + j] = (uint32_t)t; c = t >> 32; }
    r[i + NB] = (uint32_t)c;
  }
}
// r (N limbs, two's complement) += / -= x (M <= N limbs, unsigned, zero-extended)
template <int N, int M> HD_INLINE void acc_addsub(uint32_t* r, const uint32_t* x, bool subtract) {
  uint64_t c = subtract ? 1 : 0;
  for (int i = 0; i < N; i++) {
    uint32_t xi = i < M ? x[i] : 0u;
    if (subtract) xi = ~xi;
    uint64_t t = (uint64_t)r[i] + xi + c; r[i] = (uint32_t)t; c = t >> 32;
  }
}

struct GlvHalf { uint32_t mag[4]; bool neg; };

// Bandersnatch GLV split.  k: 8 canonical limbs (< r).  Lattice basis and the rounding multipliers
// G1, G2 = floor(2^352 * |n_i| / r) come from tools/gen_cuda_consts.py (asserted there to give
// |k1|,|k2| < 7/15 * 2^127, the range of 32 signed radix-16 digits).
HD_NOINLINE void band_glv_split(GlvHalf* h1, GlvHalf* h2, const uint32_t* k) {
  typedef BandConsts K;
  uint32_t g[8], prod[16], c1[5], c2[5];
  for (int i = 0; i < 8; i++) g[i] = K::G1(i);
  mul_wide<8, 8>(prod, k, g);
  { uint64_t c = (uint64_t)prod[10] + 0x80000000u; c >>= 32; for (int i = 11; i < 16; i++) { uint64_t t = (uint64_t)prod[i] + c; c1[i - 11] = (uint32_t)t; c = t >> 32; } }
  for (int i = 0; i < 8; i++) g[i] = K::G2(i);
  mul_wide<8, 8>(prod, k, g);
  { uint64_t c = (uint64_t)prod[10] + 0x80000000u; c >>= 32; for (int i = 11; i < 16; i++) { uint64_t t = (uint64_t)prod[i] + c; c2[i - 11] = (uint32_t)t; c = t >> 32; } }
  // k1 = k - c1*a1 - c2*a2 ; k2 = -c1*b1 - c2*b2   (mod 2^192, two's complement), c_i carry sign N_i_NEG
  uint32_t m[4], t[9], k1[6], k2[6];
  for (int i = 0; i < 6; i++) { k1[i] = k[i]; k2[i] = 0; }
  for (int i = 0; i < 4; i++) m[i] = K::A1(i);
  mul_wide<5, 4>(t, c1, m); acc_addsub<6, 6>(k1, t, !(K::N1_NEG ^ K::A1_NEG));
  for (int i = 0; i < 4; i++) m[i] = K::A2(i);
  mul_wide<5, 4>(t, c2, m); acc_addsub<6, 6>(k1, t, !(K::N2_NEG ^ K::A2_NEG));
  for (int i = 0; i < 4; i++) m[i] = K::B1(i);
  mul_wide<5, 4>(t, c1, m); acc_addsub<6, 6>(k2, t, !(K::N1_NEG ^ K::B1_NEG));
  for (int i = 0; i < 4; i++) m[i] = K::B2(i);
  mul_wide<5, 4>(t, c2, m); acc_addsub<6, 6>(k2, t, !(K::N2_NEG ^ K::B2_NEG));
  const uint32_t one[1] = {1};
  h1->neg = k1[5] >> 31;
  if (h1->neg) { for (int i = 0; i < 6; i++) k1[i] = ~k1[i]; acc_addsub<6, 1>(k1, one, false); }
  h2->neg = k2[5] >> 31;
  if (h2->neg) { for (int i = 0; i < 6; i++) k2[i] = ~k2[i]; acc_addsub<6, 1>(k2, one, false); }
  for (int i = 0; i < 4; i++) { h1->mag[i] = k1[i]; h2->mag[i] = k2[i]; }
}

// Signed fixed windows without a carry loop: with K' = K + 0x88..8 (resp. 0x80..80) the radix-16
// (radix-256) digit of K is nibble_i(K') - 8 (byte_i(K') - 128), in [-8,7] ([-128,127]).
// k (N limbs) += pattern repeated over the low `plimbs` limbs; limbs above only receive the carry
template <int N> HD_INLINE void add_window_bias(uint32_t* k, uint32_t pattern, int plimbs) {
  uint64_t c = 0;
  for (int i = 0; i < N; i++) { uint64_t t = (uint64_t)k[i] + (i < plimbs ? pattern : 0u) + c; k[i] = (uint32_t)t; c = t >> 32; }
}
HD_INLINE int digit4(const uint32_t* kb, int w) { return (int)((kb[w >> 3] >> ((w & 7) * 4)) & 15u) - 8; }
HD_INLINE int digit8(const uint32_t* kb, int w) { return (int)((kb[w >> 2] >> ((w & 3) * 8)) & 255u) - 128; }
HD_INLINE int digit16(const uint32_t* kb, int w) { return (int)((kb[w >> 1] >> ((w & 1) * 16)) & 65535u) - 32768; }

}  // namespace vrfs
