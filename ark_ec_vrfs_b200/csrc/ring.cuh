// the fixed columns of a ring and their KZG commitments - what `ring` -> RingContext::verifier_key /
// prover_key -> ring-proof `index` -> `PiopParams::fixed_columns` + `FixedColumns::commit` compute
// (named at /root/reference/src/lib.rs:13-17).  Three stages, all on the context's stream:
//   k_ring_columns   xs | ys | selector over the domain: keys, then the padding point up to keyset_part_size, then the caller's
//                    tail points (the powers 2^j H of the blinding base), then zero rows; selector = 1 on the keyset part
//   k_ntt_*          Radix2EvaluationDomain::ifft / fft over BLS12-381 Fr (evaluations <-> coefficients) for a monomial SRS;
//                    a Lagrange-basis SRS commits to the evaluations directly
//   msm_dev          the 3-column MSM of msm.cuh over the prepared SRS
// The row layout is a parameter list (keyset_part_size, padding, tail) rather than a constant: the ring-proof crate is not
// available offline, so the engine does not hard-code what it cannot pin ("parity unpinned" list).
#pragma once
#include "msm.cuh"
#include "gen/ntt_consts.cuh"

namespace vrfs {

typedef Fp<BlsFr> Fr255;

HD_INLINE Fr255 ntt_const(int which) {           // 0 root, 1 root_inv, 2 inv2
  Fr255 r;
  for (int i = 0; i < 8; i++) r.v[i] = which == 0 ? NttConsts::root(i) : which == 1 ? NttConsts::root_inv(i) : NttConsts::inv2(i);
  return r;
}
HD_INLINE Fr255 fr_pow_u32(Fr255 b, uint32_t e) {
  Fr255 r = Fr255::one();
  while (e) { if (e & 1u) r = r * b; b = sqr(b); e >>= 1; }
  return r;
}
// generator of the size-2^logn domain (or its inverse): TWO_ADIC_ROOT_OF_UNITY^(2^(32 - logn))
HD_INLINE Fr255 ntt_domain_gen(int logn, bool inverse) {
  Fr255 w = ntt_const(inverse ? 1 : 0);
  for (int i = logn; i < NttConsts::TWO_ADICITY; i++) w = sqr(w);
  return w;
}

// ---- BLS12-381 G1 on the wire: what ark-bls12-381's CanonicalSerialize / CanonicalDeserialize (compressed, validated) do to the
// points of a RingCommitment and of an SRS.  [RECALL, unpinned beyond the generator vector] the crate uses the zcash encoding:
// 48 bytes, big-endian x; bit 7 of byte 0 = compressed, bit 6 = infinity, bit 5 = y is the lexicographically larger root.
struct ExpG1Sqrt { template <class P> static HD_INLINE uint32_t get(int i) { return G1WireConsts::sqrt_exp(i); } };
// r = [|x|] p for the curve parameter x = -0xd201000000010000 (Hamming weight 6): MSB-first double-and-add, complete formulas
HD_INLINE void g1_mul_xabs(G1Pt& r, const G1Pt& p) {
  r = p;
#pragma unroll 1
  for (int bit = 62; bit >= 0; bit--) {
    g1_dbl(&r, &r);
    if ((G1WireConsts::X_ABS >> bit) & 1ull) sw_add<G1Curve>(&r, &r, &p);
  }
}
// prime-order subgroup test for a finite point ON THE CURVE: phi(P) = (beta x, y) equals [-x^2] P exactly on G1
// (Scott, eprint 2021/1130, sect. 6 - the test ark-bls12-381 uses; 126 doublings + 10 additions instead of a 255-bit multiplication)
HD_INLINE bool g1_in_subgroup(const Fq381& x, const Fq381& y) {
  G1Pt P, Q, Q2;
  sw_from_affine(P, x, y);
  g1_mul_xabs(Q, P);
  g1_mul_xabs(Q2, Q);                            // [x^2] P
  if (Q2.Z.is_zero()) return false;
  Fq381 beta;
  for (int i = 0; i < 12; i++) beta.v[i] = G1WireConsts::beta(i);
  return (beta * x) * Q2.Z == Q2.X && y * Q2.Z == neg(Q2.Y);
}
// 96-byte affine LE (zeros = identity) -> 48-byte compressed
HD_INLINE void g1_compress_one(uint8_t* out, const uint8_t* in) {
  uint32_t rx[12], ry[12];
  load_le<12>(rx, in); load_le<12>(ry, in + 48);
  uint32_t any = 0;
  for (int i = 0; i < 12; i++) any |= rx[i] | ry[i];
  for (int i = 0; i < 12; i++) { const uint32_t w = rx[11 - i]; out[4 * i] = (uint8_t)(w >> 24); out[4 * i + 1] = (uint8_t)(w >> 16); out[4 * i + 2] = (uint8_t)(w >> 8); out[4 * i + 3] = (uint8_t)w; }
  if (!any) { out[0] = 0xC0; return; }
  const bool high = is_high(to_mont<BlsFq>(ry));
  out[0] |= 0x80 | (high ? 0x20 : 0);
}
// 48-byte compressed -> 96-byte affine LE; false (and zeros) for a non-canonical / off-curve / out-of-subgroup encoding
HD_INLINE bool g1_decompress_one(uint8_t* out, const uint8_t* in, bool check_subgroup) {
  for (int i = 0; i < 96; i++) out[i] = 0;
  const uint8_t b0 = in[0];
  uint32_t rx[12];
  for (int i = 0; i < 12; i++) {
    const uint8_t* p = in + 4 * (11 - i);
    rx[i] = ((uint32_t)(i == 11 ? (p[0] & 0x1F) : p[0]) << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
  }
  if (!(b0 & 0x80)) return false;                // only the compressed form travels here
  uint32_t any = 0;
  for (int i = 0; i < 12; i++) any |= rx[i];
  if (b0 & 0x40) return !(b0 & 0x20) && any == 0;   // infinity: nothing else may be set
  if (!is_canonical<BlsFq>(rx)) return false;
  const Fq381 x = to_mont<BlsFq>(rx), rhs = sqr(x) * x + G1Curve::b();
  Fq381 y = pow_const<BlsFq, ExpG1Sqrt>(rhs);
  if (!(sqr(y) == rhs)) return false;            // x is not the abscissa of a point
  if (is_high(y) != ((b0 & 0x20) != 0)) y = neg(y);
  if (check_subgroup && !g1_in_subgroup(x, y)) return false;
  uint32_t ry[12];
  from_mont<BlsFq>(ry, y);
  store_le<12>(out, rx); store_le<12>(out + 48, ry);
  return true;
}

#ifdef __CUDACC__
__global__ void k_g1_compress(uint32_t n, const uint8_t* in, uint8_t* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g1_compress_one(out + (size_t)48 * i, in + (size_t)96 * i);
}
#ifndef G1_DEC_MINBLOCKS
#define G1_DEC_MINBLOCKS 4   // 128 registers: 8.3 ms for 2^17 checked points against 11.0 ms at 1 block (168 registers at 3: 8.4 ms)
#endif
__global__ void __launch_bounds__(128, G1_DEC_MINBLOCKS) k_g1_decompress(uint32_t n, const uint8_t* in, int check_subgroup, uint8_t* out, uint8_t* ok) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ok[i] = g1_decompress_one(out + (size_t)96 * i, in + (size_t)48 * i, check_subgroup != 0) ? 1 : 0;
}
// columns[0] = xs, [1] = ys, [2] = selector; every value canonical 32-byte LE (points arrive as affine x || y, 64 B)
// The kernel writes the n rows [row_lo, row_lo + n) of the domain (row_lo = 0 and n = domain size for the whole columns; a rank of
// a multi-GPU commitment passes its slice); keys points at the key of row row_lo.
__global__ void k_ring_columns(uint32_t n, uint32_t row_lo, uint32_t keyset_part, uint32_t n_keys, const uint8_t* keys, const uint8_t* padding,
                               uint32_t n_tail, const uint8_t* tail, uint8_t* columns) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t i = row_lo + j;                 // row of the domain
  const uint8_t* src = nullptr;
  if (i < n_keys) src = keys + (size_t)64 * j;
  else if (i < keyset_part) src = padding;
  else if (i - keyset_part < n_tail) src = tail + (size_t)64 * (i - keyset_part);
  uint4 z = make_uint4(0, 0, 0, 0), x0 = z, x1 = z, y0 = z, y1 = z;
  if (src) { const uint4* s = reinterpret_cast<const uint4*>(src); x0 = s[0]; x1 = s[1]; y0 = s[2]; y1 = s[3]; }
  uint4* cx = reinterpret_cast<uint4*>(columns + (size_t)32 * j);
  uint4* cy = reinterpret_cast<uint4*>(columns + (size_t)32 * ((size_t)n + j));
  uint4* cs = reinterpret_cast<uint4*>(columns + (size_t)32 * (2 * (size_t)n + j));
  cx[0] = x0; cx[1] = x1; cy[0] = y0; cy[1] = y1;
  cs[0] = make_uint4(i < keyset_part ? 1u : 0u, 0, 0, 0); cs[1] = z;
}
// the homomorphic form (ring-proof's `Ring::append`): columns[0][i] = x_i - pad_x, columns[1][i] = y_i - pad_y (mod r) on the key rows,
// zero elsewhere, so that  commit(ring) = commit(all-padding ring) + MSM(delta columns)  costs n_keys entries per window
HD_INLINE Fr255 fr_load_reduced(const uint8_t* p) {            // 16-byte aligned canonical LE -> residue < r (values < 2^256 < 3 r)
  Fr255 v;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  const uint4 lo4 = q[0], hi4 = q[1];
  v.v[0] = lo4.x; v.v[1] = lo4.y; v.v[2] = lo4.z; v.v[3] = lo4.w; v.v[4] = hi4.x; v.v[5] = hi4.y; v.v[6] = hi4.z; v.v[7] = hi4.w;
  cond_sub_p<BlsFr>(v.v, 0u); cond_sub_p<BlsFr>(v.v, 0u);
  return v;
}
HD_INLINE void fr_store16(uint8_t* p, const Fr255& v) {         // 16-byte aligned
  uint4* o = reinterpret_cast<uint4*>(p);
  o[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
  o[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}
__global__ void k_ring_delta_columns(uint32_t n, uint32_t n_keys, const uint8_t* keys, const uint8_t* padding, uint8_t* columns) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr255 dx = Fr255::zero(), dy = Fr255::zero();
  if (i < n_keys) {
    dx = fr_load_reduced(keys + (size_t)64 * i) - fr_load_reduced(padding);
    dy = fr_load_reduced(keys + (size_t)64 * i + 32) - fr_load_reduced(padding + 32);
  }
  fr_store16(columns + (size_t)32 * i, dx);
  fr_store16(columns + (size_t)32 * ((size_t)n + i), dy);
}
// tw[j] = w^j, j < n/2 (Montgomery form);
// tw[n/2] = 1/n (the scale of the inverse transform), computed once here instead of by every thread of k_ntt_store
__global__ void k_ntt_twiddles(int logn, int inverse, Fr255* tw) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x, half = (1u << logn) / 2u;
  if (j == half) tw[half] = fr_pow_u32(ntt_const(2), (uint32_t)logn);
  if (j >= half) return;
  tw[j] = fr_pow_u32(ntt_domain_gen(logn, inverse != 0), j);
}
// canonical LE -> Montgomery, bit-reversed position (values >= r are reduced, like ark-ff's from_le_bytes_mod_order)
__global__ void k_ntt_load(int logn, uint32_t ncol, const uint8_t* in, Fr255* work) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = 1u << logn;
  if (t >= (size_t)n * ncol) return;
  const uint32_t col = (uint32_t)(t >> logn), i = (uint32_t)(t & (n - 1));
  const uint4* q = reinterpret_cast<const uint4*>(in + 32 * t);
  const uint4 lo4 = q[0], hi4 = q[1];
  const uint32_t raw[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
  const uint32_t rev = logn ? __brev(i) >> (32 - logn) : 0u;
  work[(size_t)col * n + rev] = to_mont<BlsFr>(raw);
}
// stages [s0, s1) of the decimation-in-time network on bit-reversed input; one thread per butterfly per stage.
// FUSED: the block owns a contiguous span of 2^s1 elements in shared memory (s0 = 0); otherwise one stage over global memory.
#define NTT_FUSED_LOG 10
__global__ void __launch_bounds__(512) k_ntt_fused(int logn, int stages, const Fr255* tw, Fr255* work) {
  __shared__ uint4 sh_raw[(1 << NTT_FUSED_LOG) * sizeof(Fr255) / 16];
  Fr255* sh = reinterpret_cast<Fr255*>(sh_raw);
  const uint32_t span = 1u << stages, half_n = (1u << logn) / 2u;
  Fr255* base = work + (size_t)blockIdx.x * span;          // columns are contiguous, spans never straddle a column
  for (uint32_t i = threadIdx.x; i < span; i += blockDim.x) sh[i] = base[i];
  __syncthreads();
  for (int s = 1; s <= stages; s++) {
    const uint32_t half = 1u << (s - 1);
    for (uint32_t t = threadIdx.x; t < span / 2; t += blockDim.x) {
      const uint32_t k = (t >> (s - 1)) << s, j = t & (half - 1);
      const Fr255 w = tw[(size_t)j * (half_n >> (s - 1))];
      const Fr255 a = sh[k + j], b = sh[k + j + half] * w;
      sh[k + j] = a + b; sh[k + j + half] = a - b;
    }
    __syncthreads();
  }
  for (uint32_t i = threadIdx.x; i < span; i += blockDim.x) base[i] = sh[i];
}
__global__ void __launch_bounds__(256) k_ntt_stage(int logn, uint32_t ncol, int s, const Fr255* tw, Fr255* work) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = 1u << logn, half_n = n / 2u;
  if (t >= (size_t)half_n * ncol) return;
  const uint32_t col = (uint32_t)(t / half_n), u = (uint32_t)(t % half_n), half = 1u << (s - 1);
  const uint32_t k = (u >> (s - 1)) << s, j = u & (half - 1);
  Fr255* a = work + (size_t)col * n + k + j;
  const Fr255 w = tw[(size_t)j * (half_n >> (s - 1))];
  const Fr255 x = a[0], y = a[half] * w;
  a[0] = x + y; a[half] = x - y;
}
// Montgomery -> canonical LE, scaled by 1/n for the inverse transform
__global__ void k_ntt_store(int logn, uint32_t ncol, int inverse, const Fr255* work, const Fr255* tw, uint8_t* out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ((size_t)ncol << logn)) return;
  Fr255 v = work[t];
  if (inverse) v = v * tw[(1u << logn) / 2u];
  uint32_t raw[8];
  from_mont<BlsFr>(raw, v);
  uint4* o = reinterpret_cast<uint4*>(out + 32 * t);            // device buffers are 256-byte aligned: two 16-byte stores per value
  o[0] = make_uint4(raw[0], raw[1], raw[2], raw[3]);
  o[1] = make_uint4(raw[4], raw[5], raw[6], raw[7]);
}
#endif  // __CUDACC__

}  // namespace vrfs
