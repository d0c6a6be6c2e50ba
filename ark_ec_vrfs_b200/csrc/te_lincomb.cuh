// K7 + K8 of SURVEY.md 2.5 for twisted-Edwards suites: one item's linear combination
//     R = sum_j  (+/-) k_j * P_j   (variable bases, per item)   +   sum_f (+/-) s_f * B_f  (fixed bases)
// which is every group computation of the hot path:
//   ietf verify   U = s*G - c*Y            (NV=1, NF=1)      V = s*I - c*O      (NV=2)
//   ietf prove    k*G (NF=1), k*I (NV=1);  Secret::output  sk*I (NV=1);  Public  sk*G (NF=1)
//   pedersen      sk*G + b*B (NF=2), k*G + kb*B (NF=2), k*I, s*I - c*O (NV=2), c*Yb - s*G - sb*B (NV=1, NF=2)
// The reference does each k*P with ark-ec's bit-serial `mul_bigint`; the result is the same group
// element, and every consumer works on its canonical affine encoding.
//
// Variable bases: signed radix-16 fixed windows (no data-dependent branches -> no warp divergence),
// Bandersnatch scalars GLV-split into two 127-bit halves acting on P and psi(P) so a K-term sum costs
// 124 shared doublings + 32*2K additions.  Window tables (9 cached entries per base incl. the identity)
// live in a per-thread slab of global memory (L2-resident across the resident grid), read with 16-byte loads.
// Fixed bases: signed radix-256 windows over a precomputed affine table (32 x 129 entries, 396 KB, L2-resident).
#pragma once
#include "te.cuh"
#include "scalar.cuh"

namespace vrfs {

#if !defined(__CUDACC__)
struct uint4 { uint32_t x, y, z, w; };
#endif

template <class C> struct TeTraits {
  static constexpr int SPLIT = C::HAS_GLV ? 2 : 1;          // tables per variable base
  static constexpr int WINDOWS = C::HAS_GLV ? 32 : 64;      // radix-16 windows per (half-)scalar
  static constexpr int KB_LIMBS = C::HAS_GLV ? 4 : 8;
  static constexpr int TBL_ENTRIES = 9;                     // 0*P .. 8*P
  static constexpr int FIX_WINDOWS = 32, FIX_ENTRIES = 129; // radix-256, 0..128
};
// bytes of per-thread table slab for NV variable bases
#if defined(__CUDACC__)
#define VRFS_HD __host__ __device__
#else
#define VRFS_HD
#endif
template <class C> VRFS_HD constexpr size_t te_slab_bytes(int nv) { return (size_t)nv * TeTraits<C>::SPLIT * TeTraits<C>::TBL_ENTRIES * sizeof(TECached<C>); }

struct VarTerm { const uint8_t* pts; uint32_t pt_stride; const uint8_t* sc; uint32_t sc_stride; uint32_t negate; };
struct FixTerm { const uint8_t* sc; uint32_t sc_stride; uint32_t negate; const void* table; };
struct LincombArgs {
  uint32_t n;
  VarTerm var[2];
  FixTerm fix[2];
  uint32_t* out_xyz;     // n x 24 limbs: X, Y, Z (Montgomery form)
  uint8_t* valid;        // n bytes, AND-ed with "all variable bases canonical and on the curve" (may be null)
  uint8_t* slab;         // per-thread table slabs
};

template <class T> HD_INLINE void copy_words16(T* dst, const T* src_) {  // sizeof(T) % 16 == 0, both 16-byte aligned
  const uint4* s = reinterpret_cast<const uint4*>(src_);
  uint4* d = reinterpret_cast<uint4*>(dst);
  for (unsigned i = 0; i < sizeof(T) / 16; i++) d[i] = s[i];
}

// affine x||y, 32-byte little-endian canonical each -> Montgomery; false if not canonical / not on the curve
template <class C> HD_INLINE bool te_load_affine(typename C::F& x, typename C::F& y, const uint8_t* p) {
  uint32_t rx[8], ry[8];
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1], c = q[2], d = q[3];
  rx[0] = a.x; rx[1] = a.y; rx[2] = a.z; rx[3] = a.w; rx[4] = b.x; rx[5] = b.y; rx[6] = b.z; rx[7] = b.w;
  ry[0] = c.x; ry[1] = c.y; ry[2] = c.z; ry[3] = c.w; ry[4] = d.x; ry[5] = d.y; ry[6] = d.z; ry[7] = d.w;
  bool ok = is_canonical<typename C::Fq>(rx) & is_canonical<typename C::Fq>(ry);
  x = to_mont<typename C::Fq>(rx); y = to_mont<typename C::Fq>(ry);
  return ok & te_on_curve<C>(x, y);
}
// 32-byte little-endian scalar, reduced mod r (codec scalar_decode = from_le_bytes_mod_order) -> canonical limbs
template <class C> HD_INLINE void load_scalar_mod_r(uint32_t* k, const uint8_t* p) {
  uint32_t raw[8];
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w; raw[4] = b.x; raw[5] = b.y; raw[6] = b.z; raw[7] = b.w;
  Fp<typename C::Fr> m = to_mont<typename C::Fr>(raw);
  from_mont<typename C::Fr>(k, m);
}

// table of j*B, j = 0..8, in cached form; B given in extended coordinates
template <class C> HD_INLINE void te_build_table(TECached<C>* tbl, const TEPoint<C>& B) {
  TEPoint<C> cur; te_set_identity(cur);
  TECached<C> cb, e;
  te_to_cached(cb, B);
  te_to_cached(e, cur); copy_words16(&tbl[0], &e);
  copy_words16(&tbl[1], &cb);
  cur = B;
#pragma unroll 1
  for (int j = 2; j <= 8; j++) {
    te_add_cached<C>(&cur, &cur, &cb, false);
    te_to_cached(e, cur); copy_words16(&tbl[j], &e);
  }
}

template <class C, int NV, int NF>
HD_INLINE bool te_lincomb_item(const LincombArgs& A, uint32_t item, TECached<C>* slab, TEPoint<C>& acc) {
  typedef TeTraits<C> T;
  typedef typename C::F F;
  constexpr int NT = NV * T::SPLIT;
  bool ok = true;
  te_set_identity(acc);
  if (NV > 0) {
    uint32_t kb[NT > 0 ? NT : 1][T::KB_LIMBS];
    bool kneg[NT > 0 ? NT : 1];
#pragma unroll
    for (int v = 0; v < NV; v++) {
      F x, y;
      ok &= te_load_affine<C>(x, y, A.var[v].pts + (size_t)item * A.var[v].pt_stride);
      TEPoint<C> B; te_from_affine<C>(B, x, y);
      uint32_t k[8];
      load_scalar_mod_r<C>(k, A.var[v].sc + (size_t)item * A.var[v].sc_stride);
      bool neg = A.var[v].negate != 0;
      if constexpr (C::HAS_GLV) {
        GlvHalf h1, h2;
        band_glv_split(&h1, &h2, k);
        for (int i = 0; i < 4; i++) { kb[T::SPLIT * v][i] = h1.mag[i]; kb[T::SPLIT * v + T::SPLIT - 1][i] = h2.mag[i]; }
        kneg[T::SPLIT * v] = h1.neg ^ neg; kneg[T::SPLIT * v + T::SPLIT - 1] = h2.neg ^ neg;
        te_build_table<C>(slab + (T::SPLIT * v) * T::TBL_ENTRIES, B);
        TEPoint<C> E;
        band_endo(reinterpret_cast<TEPoint<BandCurve>*>(&E), reinterpret_cast<const TEPoint<BandCurve>*>(&B));
        te_build_table<C>(slab + (T::SPLIT * v + T::SPLIT - 1) * T::TBL_ENTRIES, E);
      } else {
        for (int i = 0; i < 8; i++) kb[v][i % T::KB_LIMBS] = k[i];
        kneg[v] = neg;
        te_build_table<C>(slab + v * T::TBL_ENTRIES, B);
      }
    }
    for (int t = 0; t < NT; t++) add_window_bias<T::KB_LIMBS>(kb[t], 0x88888888u);
#pragma unroll 1
    for (int w = T::WINDOWS - 1; w >= 0; w--) {
      if (w != T::WINDOWS - 1) {
        te_dbl<C>(&acc, &acc, false); te_dbl<C>(&acc, &acc, false); te_dbl<C>(&acc, &acc, false); te_dbl<C>(&acc, &acc, true);
      }
#pragma unroll
      for (int t = 0; t < NT; t++) {
        int d = digit4(kb[t], w);
        int idx = d < 0 ? -d : d;
        TECached<C> e;
        copy_words16(&e, &slab[t * T::TBL_ENTRIES + idx]);
        te_add_cached<C>(&acc, &acc, &e, (d < 0) != kneg[t]);
      }
    }
  }
#pragma unroll
  for (int f = 0; f < NF; f++) {
    uint32_t k[8];
    load_scalar_mod_r<C>(k, A.fix[f].sc + (size_t)item * A.fix[f].sc_stride);
    add_window_bias<8>(k, 0x80808080u);
    const TEAffCached<C>* tbl = reinterpret_cast<const TEAffCached<C>*>(A.fix[f].table);
    bool neg = A.fix[f].negate != 0;
#pragma unroll 1
    for (int w = 0; w < T::FIX_WINDOWS; w++) {
      int d = digit8(k, w);
      int idx = d < 0 ? -d : d;
      TEAffCached<C> e;
      copy_words16(&e, &tbl[w * T::FIX_ENTRIES + idx]);
      te_madd<C>(&acc, &acc, &e, (d < 0) != neg);
    }
  }
  return ok;
}

// one fixed-base table entry: (d * 256^w) * B, affine cached.  Used once per context by the table kernel.
template <class C>
HD_INLINE void te_fixed_table_entry(TEAffCached<C>& out, const typename C::F& bx, const typename C::F& by, int w, int d) {
  TEPoint<C> P, R; te_from_affine<C>(P, bx, by);
  for (int i = 0; i < 8 * w; i++) te_dbl<C>(&P, &P, true);
  te_set_identity(R);
  for (int bit = 7; bit >= 0; bit--) {
    te_dbl<C>(&R, &R, true);
    if ((d >> bit) & 1) te_add<C>(&R, &R, &P);
  }
  typename C::F zi = inv(R.Z);
  out.x = R.X * zi; out.y = R.Y * zi; out.dt = out.x * out.y * C::d();
}

}  // namespace vrfs
