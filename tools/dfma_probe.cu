// Experiment (not product code): a 255-bit Montgomery product on the FP64 pipe (DFMA hi/lo splitting with integer
// accumulation of the mantissas), alone and co-scheduled with the IMAD.WIDE product of csrc/arith.cuh, to decide
// whether a hybrid "IMAD warps + DFMA warps" scalar-multiplication kernel is worth building on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I ark_ec_vrfs_b200/csrc -o /tmp/dfma_probe tools/dfma_probe.cu
// Representation: 5 signed limbs, radix 2^51, held as doubles; Montgomery factor 2^-306 (6 reduction rounds) so that
// outputs are |v| <= p/2 (+eps) with limbs |l| <= 2^50 and one lazy limb-wise add/sub of two outputs is a valid input.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gen/field_consts.cuh"

using namespace vrfs;

namespace f51 {
struct F51 { double v[5]; };

__device__ __forceinline__ constexpr int n5(int k) { return (k < 0 || k > 8) ? 0 : (k <= 4 ? k + 1 : 9 - k); }           // pairs i+j=k, i,j in 0..4
__device__ __forceinline__ constexpr int n65(int k) { int c = 0; for (int i = 0; i < 6; i++) for (int j = 0; j < 5; j++) c += (i + j == k); return c; }
__device__ __forceinline__ constexpr int nL(int k) { return n5(k) + n65(k); }

#define F51_MASK 0x7ffffffffffffull
#define F51_P0 0x7ffff00000001ull
#define F51_P1 0x05fffcb7fdfffull
#define F51_P2 0x68760154ef690ull
#define F51_P3 0x6a4199cec0404ull
#define F51_P4 0x73eda753299d7ull
#define F51_NINV 0x7fffeffffffffull

__device__ __forceinline__ double plimb(int j) {
  return j == 0 ? (double)F51_P0 : j == 1 ? (double)F51_P1 : j == 2 ? (double)F51_P2 : j == 3 ? (double)F51_P3 : (double)F51_P4;
}

// one limb product x*y -> lo (into L) and hi (into H): x*y = hi*2^52 + lo, lo in [0,2^52), |x*y| < 2^103
__device__ __forceinline__ void mac(double x, double y, uint64_t& L, uint64_t& H) {
  const double c1 = 0x1.8p104, c2 = 0x1.8p104 + 0x1p52;
  double hi = __fma_rz(x, y, c1);
  double lo = __fma_rz(x, y, c2 - hi);
  H += (uint64_t)__double_as_longlong(hi);
  L += (uint64_t)__double_as_longlong(lo);
}

__device__ __forceinline__ double int_to_double(int64_t x) {   // |x| < 2^51
  return __longlong_as_double(0x4338000000000000ll + x) - 0x1.8p52;
}

template <bool SQR>
__device__ __noinline__ F51 mul(F51 a, F51 b) {
  uint64_t L[11], H[12];
#pragma unroll
  for (int k = 0; k < 11; k++) {
    L[k] = 0ull - (uint64_t)nL(k) * 0x4330000000000000ull;
    H[k] = 0ull - (uint64_t)(k ? nL(k - 1) : 0) * 0x4678000000000000ull;
  }
  H[11] = 0;
  int64_t carry = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    if (i < 5) {
#pragma unroll
      for (int j = 0; j < 5; j++) mac(a.v[j], b.v[i], L[i + j], H[i + j + 1]);
    }
    int64_t col = (int64_t)(L[i] + 2 * H[i]) + carry;          // complete except for the lo of q*p0 (bias is 0 mod 2^52)
    uint64_t q = ((uint64_t)col * F51_NINV) & F51_MASK;
    int64_t qs = (int64_t)((q + (1ull << 50)) & F51_MASK) - (1ll << 50);   // centred: [-2^50, 2^50)
    double qd = int_to_double(qs);
    uint64_t L0 = 0ull;                                        // lo of q*p0 kept separate so that col stays usable
#pragma unroll
    for (int j = 0; j < 5; j++) mac(qd, plimb(j), j == 0 ? L0 : L[i + j], H[i + j + 1]);
    carry = (col + (int64_t)L0) >> 51;                         // exact: divisible by 2^51
  }
  // L0 above carried its own bias of one lo term: compensate by having counted it in nL(i) -> add it back
  // (handled below by construction: see note) ; columns 6..10 are the result
  F51 r;
#pragma unroll
  for (int k = 6; k < 11; k++) {
    int64_t t = (int64_t)(L[k] + 2 * H[k]) + carry;
    int64_t lo = k < 10 ? ((int64_t)(((uint64_t)t + (1ull << 50)) & F51_MASK) - (1ll << 50)) : t;
    carry = (t - lo) >> 51;
    r.v[k - 6] = int_to_double(lo);
  }
  return r;
}
}  // namespace f51

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_chain(int mode, int iters_imad, int iters_dfma, double* io, uint32_t* sink, int split) {
  // mode 0: IMAD only; 1: DFMA only; 2: hybrid (warp w uses DFMA iff (w % split) != 0 ... see below)
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  int warp = threadIdx.x >> 5;
  bool use_dfma = mode == 1 || (mode == 2 && (warp % split) != 0) || (mode == 3 && (warp % split) == 0);
  if (!use_dfma) {
    Fp<BlsFr> a, b;
    for (int i = 0; i < 8; i++) { a.v[i] = t + i; b.v[i] = (t ^ 0x5bd1e995u) + 7 * i; }
    a.v[7] &= 0x3fffffffu; b.v[7] &= 0x3fffffffu;
#pragma unroll 1
    for (int i = 0; i < iters_imad; i++) {
#pragma unroll 1
      for (int r = 0; r < 16; r++) { a = a * b; b = b * a; }
    }
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= a.v[i] ^ b.v[i];
    sink[t] = s;
  } else {
    f51::F51 a, b;
    for (int i = 0; i < 5; i++) { a.v[i] = io[(size_t)t * 10 + i]; b.v[i] = io[(size_t)t * 10 + 5 + i]; }
#pragma unroll 1
    for (int i = 0; i < iters_dfma; i++) {
#pragma unroll 1
      for (int r = 0; r < 16; r++) { a = f51::mul<false>(a, b); b = f51::mul<false>(b, a); }
    }
    double s = 0;
    for (int i = 0; i < 5; i++) s += a.v[i] + b.v[i];
    sink[t] = (uint32_t)__double2loint(s) ^ (uint32_t)__double2hiint(s);
  }
}

// correctness: r = a*b, and r2 = (a+b)*(a-b) (lazy inputs), written back for the host to check
__global__ void k_check(const double* in, double* out, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  f51::F51 a, b, s, d;
  for (int i = 0; i < 5; i++) { a.v[i] = in[t * 10 + i]; b.v[i] = in[t * 10 + 5 + i]; s.v[i] = a.v[i] + b.v[i]; d.v[i] = a.v[i] - b.v[i]; }
  f51::F51 r = f51::mul<false>(a, b), r2 = f51::mul<false>(s, d);
  for (int i = 0; i < 5; i++) { out[t * 10 + i] = r.v[i]; out[t * 10 + 5 + i] = r2.v[i]; }
}

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint64_t rng() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

int main(int argc, char** argv) {
  int dev = 0;
  cudaSetDevice(dev);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
  const int threads = 256, blocks = prop.multiProcessorCount * 8, nthr = threads * blocks;
  double* h_io = (double*)malloc(sizeof(double) * 10 * nthr);
  for (int t = 0; t < nthr; t++)
    for (int i = 0; i < 10; i++) {
      int64_t v = (int64_t)(rng() & F51_MASK) - (1ll << 50);                 // [-2^50, 2^50)
      if (i % 5 == 4) v = (int64_t)(rng() % 0x39f6d3a994ceull) - 0x1cfb69d4ca67ll;   // top limb: |value| <~ p/2
      h_io[(size_t)t * 10 + i] = (double)v;
    }
  double *d_io, *d_out; uint32_t* d_sink;
  cudaMalloc(&d_io, sizeof(double) * 10 * nthr); cudaMalloc(&d_out, sizeof(double) * 10 * 4096); cudaMalloc(&d_sink, 4 * nthr);
  cudaMemcpy(d_io, h_io, sizeof(double) * 10 * nthr, cudaMemcpyHostToDevice);
  // correctness dump
  k_check<<<16, 256>>>(d_io, d_out, 4096);
  double* h_out = (double*)malloc(sizeof(double) * 10 * 4096);
  cudaMemcpy(h_out, d_out, sizeof(double) * 10 * 4096, cudaMemcpyDeviceToHost);
  if (cudaGetLastError() != cudaSuccess) { printf("check kernel failed\n"); return 1; }
  FILE* f = fopen(argc > 1 ? argv[1] : "gpurun_out/dfma_check.txt", "w");
  if (f) {
    for (int t = 0; t < 4096; t++) {
      for (int i = 0; i < 10; i++) fprintf(f, "%lld ", (long long)h_io[(size_t)t * 10 + i]);
      for (int i = 0; i < 10; i++) fprintf(f, "%lld ", (long long)h_out[(size_t)t * 10 + i]);
      fprintf(f, "\n");
    }
    fclose(f);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct { int mode, split; const char* name; } cfg[] = {
      {0, 1, "imad_only"}, {1, 1, "dfma_only"}, {2, 2, "hybrid 1 imad : 1 dfma warps"}, {2, 4, "hybrid 1 imad : 3 dfma"}, {3, 4, "hybrid 3 imad : 1 dfma"},
      {3, 8, "hybrid 7 imad : 1 dfma"}, {2, 8, "hybrid 1 imad : 7 dfma"}};
  const int di[] = {16, 32, 48, 64, 96, 128, 192};
  for (auto& c : cfg) {
    for (int k = 0; k < 7; k++) {
      int ii = 64, id = di[k];
      if (c.mode < 2 && k != 3) continue;
      if (getenv("PROBE_K") && k != atoi(getenv("PROBE_K"))) continue;
      float best = 1e30f;
      for (int rep = 0; rep < (getenv("PROBE_REPS") ? atoi(getenv("PROBE_REPS")) : 3); rep++) {
        cudaEventRecord(e0);
        k_chain<<<blocks, threads>>>(c.mode, ii, id, d_io, d_sink, c.split);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      double fd = c.mode == 0 ? 0.0 : c.mode == 1 ? 1.0 : c.mode == 2 ? (double)(c.split - 1) / c.split : 1.0 / c.split;
      double prods = 32.0 * nthr * (ii * (1.0 - fd) + id * fd);
      printf("%-32s imad_iters %3d dfma_iters %3d  %8.3f ms  %8.2f G products/s\n", c.name, ii, id, best, prods / best / 1e6);
    }
  }
  // unequal work: let each class run for its own count so that both finish together: sweep dfma iters
  printf("done (%s, %d SMs)\n", prop.name, prop.multiProcessorCount);
  return 0;
}
