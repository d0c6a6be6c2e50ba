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

#ifdef __CUDACC__
// columns[0] = xs, [1] = ys, [2] = selector; every value canonical 32-byte LE (points arrive as affine x || y, 64 B)
__global__ void k_ring_columns(uint32_t n, uint32_t keyset_part, uint32_t n_keys, const uint8_t* keys, const uint8_t* padding,
                               uint32_t n_tail, const uint8_t* tail, uint8_t* columns) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* src = nullptr;
  if (i < n_keys) src = keys + (size_t)64 * i;
  else if (i < keyset_part) src = padding;
  else if (i - keyset_part < n_tail) src = tail + (size_t)64 * (i - keyset_part);
  uint4 z = make_uint4(0, 0, 0, 0), x0 = z, x1 = z, y0 = z, y1 = z;
  if (src) { const uint4* s = reinterpret_cast<const uint4*>(src); x0 = s[0]; x1 = s[1]; y0 = s[2]; y1 = s[3]; }
  uint4* cx = reinterpret_cast<uint4*>(columns + (size_t)32 * i);
  uint4* cy = reinterpret_cast<uint4*>(columns + (size_t)32 * ((size_t)n + i));
  uint4* cs = reinterpret_cast<uint4*>(columns + (size_t)32 * (2 * (size_t)n + i));
  cx[0] = x0; cx[1] = x1; cy[0] = y0; cy[1] = y1;
  cs[0] = make_uint4(i < keyset_part ? 1u : 0u, 0, 0, 0); cs[1] = z;
}
// the homomorphic form (ring-proof's `Ring::append`): columns[0][i] = x_i - pad_x, columns[1][i] = y_i - pad_y (mod r) on the key rows,
// zero elsewhere, so that  commit(ring) = commit(all-padding ring) + MSM(delta columns)  costs n_keys entries per window
HD_INLINE Fr255 fr_load_reduced(const uint8_t* p) {            // canonical LE -> residue < r (values < 2^256 < 3 r)
  Fr255 v;
  load_le<8>(v.v, p);
  cond_sub_p<BlsFr>(v.v, 0u); cond_sub_p<BlsFr>(v.v, 0u);
  return v;
}
__global__ void k_ring_delta_columns(uint32_t n, uint32_t n_keys, const uint8_t* keys, const uint8_t* padding, uint8_t* columns) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr255 dx = Fr255::zero(), dy = Fr255::zero();
  if (i < n_keys) {
    dx = fr_load_reduced(keys + (size_t)64 * i) - fr_load_reduced(padding);
    dy = fr_load_reduced(keys + (size_t)64 * i + 32) - fr_load_reduced(padding + 32);
  }
  store_le<8>(columns + (size_t)32 * i, dx.v);
  store_le<8>(columns + (size_t)32 * ((size_t)n + i), dy.v);
}
// tw[j] = w^j, j < n/2 (Montgomery form)
__global__ void k_ntt_twiddles(int logn, int inverse, Fr255* tw) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= (1u << logn) / 2u) return;
  tw[j] = fr_pow_u32(ntt_domain_gen(logn, inverse != 0), j);
}
// canonical LE -> Montgomery, bit-reversed position (values >= r are reduced, like ark-ff's from_le_bytes_mod_order)
__global__ void k_ntt_load(int logn, uint32_t ncol, const uint8_t* in, Fr255* work) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = 1u << logn;
  if (t >= (size_t)n * ncol) return;
  const uint32_t col = (uint32_t)(t >> logn), i = (uint32_t)(t & (n - 1));
  uint32_t raw[8];
  load_le<8>(raw, in + 32 * t);
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
__global__ void k_ntt_store(int logn, uint32_t ncol, int inverse, const Fr255* work, uint8_t* out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ((size_t)ncol << logn)) return;
  Fr255 v = work[t];
  if (inverse) v = v * fr_pow_u32(ntt_const(2), (uint32_t)logn);
  uint32_t raw[8];
  from_mont<BlsFr>(raw, v);
  store_le<8>(out + 32 * t, raw);
}
#endif  // __CUDACC__

}  // namespace vrfs
