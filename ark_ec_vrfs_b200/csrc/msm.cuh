// K12 of SURVEY.md 2.5: Pippenger bucket MSM over BLS12-381 G1 for the ring-VRF KZG commitment
// (`ring` -> ring-proof `index`/`commit` -> ark-ec `VariableBaseMSM::msm`, named at
// /root/reference/src/lib.rs:13-17): n_columns scalar columns over one base vector.
// The result is a group element; its canonical affine form does not depend on how the sum is grouped, so
// the window size, signed digits and bucket order used here need not match ark-ec's.
//
// Pipeline (all on the context's stream):
//   k_msm_prep_bases   canonical LE affine -> Montgomery affine (identity = all zero)
//   k_msm_histogram    signed radix-2^c digits of every scalar; per-(column, window, bucket) counts
//   k_msm_scan         exclusive scan of the counts inside each (column, window) segment
//   k_msm_scatter      point indices (+ sign bit) grouped by bucket
//   k_msm_accumulate   one thread per bucket: complete additions of its points
//   k_msm_window       one block per (column, window): chunked running sums  sum_j j*B_j, shared-memory tree
//   k_msm_final        one thread per column: Horner over the windows; projective partial or affine out
#pragma once
#include "lincomb.cuh"

namespace vrfs {

typedef SWPoint<G1Curve> G1Pt;
typedef Fp<BlsFq> Fq381;
struct G1Aff { Fq381 x, y; };   // Montgomery form; x = y = 0 encodes the identity ((0,0) is not on the curve)

struct MsmPlan {
  uint32_t n, ncol;
  int c, windows, nb;            // window bits, window count, buckets per window (2^(c-1))
};
VRFS_HD inline MsmPlan msm_plan(uint32_t n, uint32_t ncol) {
  MsmPlan p; p.n = n; p.ncol = ncol;
  int lg = 0; while ((1u << (lg + 1)) <= n) lg++;
  p.c = lg <= 9 ? 7 : lg <= 11 ? 9 : lg <= 13 ? 11 : lg <= 15 ? 12 : 13;
  p.windows = (255 + p.c) / p.c;          // 255 scalar bits + the top carry of the signed recoding
  p.nb = 1 << (p.c - 1);
  return p;
}
#define MSM_CHUNK 16   // buckets per thread in the window reduction

HD_INLINE void g1_load_aff(G1Pt& P, const G1Aff* a, bool negate) {
  uint4* d = reinterpret_cast<uint4*>(&P);
  const uint4* s = reinterpret_cast<const uint4*>(a);
  for (int i = 0; i < 6; i++) d[i] = s[i];
  bool inf = P.X.is_zero() & P.Y.is_zero();
  P.Y = cneg(P.Y, negate);
  P.Z = select(inf, Fq381::zero(), Fq381::one());
  P.Y = select(inf, Fq381::one(), P.Y);
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
  from_mont<BlsFr>(k, to_mont<BlsFr>(raw));     // reduce mod r like the reference's scalar decode
}

#ifdef __CUDACC__
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
// counts[(col*W + w)*nb + (|d|-1)]++
__global__ void __launch_bounds__(128) k_msm_histogram(MsmPlan p, const uint8_t* scalars, uint32_t* counts) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.n * p.ncol) return;
  uint32_t col = t / p.n;
  uint32_t k[8];
  msm_load_scalar(k, scalars + (size_t)32 * t);
  int carry = 0;
  for (int w = 0; w < p.windows; w++) {
    int d = msm_digit(k, w, p.c, carry);
    if (d != 0) atomicAdd(&counts[((size_t)col * p.windows + w) * p.nb + (d < 0 ? -d : d) - 1], 1u);
  }
}
// one block per (col, window): exclusive scan of nb counts -> offsets (relative to the segment), cursors zeroed
__global__ void __launch_bounds__(256) k_msm_scan(MsmPlan p, const uint32_t* counts, uint32_t* offsets) {
  __shared__ uint32_t part[256];
  const uint32_t seg = blockIdx.x;
  const uint32_t* c = counts + (size_t)seg * p.nb;
  uint32_t* o = offsets + (size_t)seg * p.nb;
  const int per = (p.nb + 255) / 256;
  const int lo = threadIdx.x * per, hi = min(lo + per, p.nb);
  uint32_t s = 0;
  for (int i = lo; i < hi; i++) s += c[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t run = 0; for (int i = 0; i < 256; i++) { uint32_t v = part[i]; part[i] = run; run += v; } }
  __syncthreads();
  uint32_t run = part[threadIdx.x];
  for (int i = lo; i < hi; i++) { o[i] = run; run += c[i]; }
}
// list[seg*n + offsets[bucket] + pos] = i | sign << 31
__global__ void __launch_bounds__(128) k_msm_scatter(MsmPlan p, const uint8_t* scalars, const uint32_t* offsets, uint32_t* cursors, uint32_t* list) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.n * p.ncol) return;
  uint32_t col = t / p.n, i = t % p.n;
  uint32_t k[8];
  msm_load_scalar(k, scalars + (size_t)32 * t);
  int carry = 0;
  for (int w = 0; w < p.windows; w++) {
    int d = msm_digit(k, w, p.c, carry);
    if (d == 0) continue;
    size_t seg = (size_t)col * p.windows + w;
    size_t b = seg * p.nb + (d < 0 ? -d : d) - 1;
    uint32_t pos = atomicAdd(&cursors[b], 1u);
    list[seg * p.n + offsets[b] + pos] = i | (d < 0 ? 0x80000000u : 0u);
  }
}
// one thread per bucket
__global__ void __launch_bounds__(128) k_msm_accumulate(MsmPlan p, const G1Aff* bases, const uint32_t* counts, const uint32_t* offsets,
                                                         const uint32_t* list, G1Pt* buckets) {
  size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)p.ncol * p.windows * p.nb;
  if (b >= total) return;
  size_t seg = b / p.nb;
  const uint32_t* l = list + seg * p.n + offsets[b];
  uint32_t cnt = counts[b];
  G1Pt acc; sw_set_identity(acc);
  for (uint32_t j = 0; j < cnt; j++) {
    uint32_t e = l[j];
    G1Pt q;
    g1_load_aff(q, &bases[e & 0x7fffffffu], (e >> 31) != 0);
    sw_add<G1Curve>(&acc, &acc, &q);
  }
  buckets[b] = acc;
}
// r = k * p for a small non-negative k (double-and-add, MSB first)
__device__ __forceinline__ void g1_mul_small(G1Pt& r, const G1Pt& p, uint32_t k) {
  sw_set_identity(r);
  for (int bit = 31 - __clz(k | 1u); bit >= 0; bit--) {
    sw_add<G1Curve>(&r, &r, &r);
    if ((k >> bit) & 1u) sw_add<G1Curve>(&r, &r, &p);
  }
}
// one block per (col, window), nb / MSM_CHUNK threads: window sum = sum_{j=1..nb} j * B_j
__global__ void k_msm_window(MsmPlan p, const G1Pt* buckets, G1Pt* window_sums) {
  extern __shared__ uint4 smem_raw[];
  G1Pt* sh = reinterpret_cast<G1Pt*>(smem_raw);
  const uint32_t seg = blockIdx.x, t = threadIdx.x;
  const G1Pt* B = buckets + (size_t)seg * p.nb;
  const int lo = t * MSM_CHUNK;            // bucket index j-1 in [lo, lo + CHUNK)
  G1Pt run, tot; sw_set_identity(run); sw_set_identity(tot);
  for (int j = lo + MSM_CHUNK - 1; j >= lo; j--) {
    G1Pt q = B[j];
    sw_add<G1Curve>(&run, &run, &q);
    sw_add<G1Curve>(&tot, &tot, &run);
  }
  // tot = sum (j - lo + 1) * B_j (bucket value j+1 at index j)  ->  add lo * run
  if (lo > 0) { G1Pt m; g1_mul_small(m, run, (uint32_t)lo); sw_add<G1Curve>(&tot, &tot, &m); }
  sh[t] = tot;
  __syncthreads();
  for (int stride = blockDim.x >> 1; stride > 0; stride >>= 1) {
    if ((int)t < stride) { G1Pt a = sh[t], b = sh[t + stride]; sw_add<G1Curve>(&a, &a, &b); sh[t] = a; }
    __syncthreads();
  }
  if (t == 0) window_sums[seg] = sh[0];
}
// Horner over windows; out_mode 0: affine LE canonical (96 B, identity = zeros); 1: projective X,Y,Z LE canonical (144 B)
__global__ void k_msm_final(MsmPlan p, const G1Pt* window_sums, uint8_t* out, int out_mode) {
  uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= p.ncol) return;
  const G1Pt* W = window_sums + (size_t)col * p.windows;
  G1Pt acc = W[p.windows - 1];
  for (int w = p.windows - 2; w >= 0; w--) {
    for (int k = 0; k < p.c; k++) sw_add<G1Curve>(&acc, &acc, &acc);
    G1Pt q = W[w];
    sw_add<G1Curve>(&acc, &acc, &q);
  }
  uint32_t raw[12];
  if (out_mode == 1) {
    uint8_t* o = out + (size_t)144 * col;
    from_mont<BlsFq>(raw, acc.X); store_le<12>(o, raw);
    from_mont<BlsFq>(raw, acc.Y); store_le<12>(o + 48, raw);
    from_mont<BlsFq>(raw, acc.Z); store_le<12>(o + 96, raw);
  } else {
    uint8_t* o = out + (size_t)96 * col;
    Fq381 zi = inv(acc.Z);                      // identity: Z = 0 -> zi = 0 -> zeros
    from_mont<BlsFq>(raw, acc.X * zi); store_le<12>(o, raw);
    from_mont<BlsFq>(raw, acc.Y * zi); store_le<12>(o + 48, raw);
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
  Fq381 zi = inv(acc.Z);
  from_mont<BlsFq>(raw, acc.X * zi); store_le<12>(o, raw);
  from_mont<BlsFq>(raw, acc.Y * zi); store_le<12>(o + 48, raw);
}
#endif  // __CUDACC__

}  // namespace vrfs
