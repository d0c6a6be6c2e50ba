// SURVEY.md 8(f)3, second form: ONE product of two BLS12-381 pairings spread over the lanes of one or two warps.
//
// The one-thread form (csrc/pairing.cuh) needs 27 ms of a B200 for the product that ends a batched KZG check
// (`Bls12::multi_miller_loop` + `final_exponentiation`, named through /root/reference/src/lib.rs:13-17 `ring`): ~19 000 F_q
// products in one dependency chain as far as a single thread can tell.  At the F_q level the work is wide - an F_q12 product is
// 54 independent F_q products between thin layers of additions - so tools/gen_pairing_prog.py traces the very formulas of
// pairing.cuh into a DAG, schedules it into STEPS of up to 32 independent operations and allocates the values to a register
// file of F_q elements.  What runs here is the interpreter of those tables ("lane programs"):
//   * register file: PairingProg::NREG F_q elements (48 B, Montgomery form) in shared memory, one block of PairingProg::LANES threads per product;
//   * one step = every lane decodes its 128-bit operation (a product of two registers, or a SIGNED SUM of up to eight registers, each times +-1..8 -
//     the generator flattens the towers' chains of additions into such sums, 2 800 steps instead of 6 900), reads its
//     operands, computes, writes dst, __syncwarp();
//     the generator guarantees that no register written in a step is read by ANOTHER lane in that step, so the one barrier per
//     step is the only synchronisation;
//   * the MACRO program strings the segments together (63 Miller iterations, easy part, five exponentiations by the curve
//     parameter as 63 cyclotomic squarings + 5 products each, the glue products).
// The generator executes the emitted tables in a Python model and checks them against the one-thread twin and the naive oracle
// before writing them; tests/test_pairing_host.py runs THIS interpreter on the host against the oracle, tests/test_gpu_pairing.py
// on the GPU.  Pairs with an identity point (which contribute 1 and are skipped by the reference) and malformed inputs take the
// one-thread path - they are the rare case.
#pragma once
#include "pairing.cuh"
#include "gen/pairing_prog.cuh"

namespace vrfs {

// a / 2 mod q (Montgomery form is linear: (aR)/2 = (a/2)R)
HD_INLINE Fq381 fq381_half(const Fq381& a) {
  uint32_t t[12], pm[12];
  const uint32_t odd = a.v[0] & 1u;
  for (int i = 0; i < 12; i++) pm[i] = odd ? BlsFq::mod(i) : 0u;
  MontChains<12>::add(t, a.v, pm);                 // < 2q < 2^382: no carry out
  Fq381 r;
  for (int i = 0; i < 11; i++) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
  r.v[11] = t[11] >> 1;
  return r;
}

// ---- one operation of one lane on the register file (format: gen/pairing_prog.cuh) ---------------------------------------------
struct PairingOp { uint32_t w[4]; };
HD_INLINE uint32_t pairing_op_field(const PairingOp& op, int j) {          // source j: register | sign << 9 | (multiplier - 1) << 10
  const int off = 16 + 13 * j, wd = off >> 5, sh = off & 31;
  uint32_t v = op.w[wd] >> sh;
  if (sh > 19) v |= op.w[wd + 1] << (32 - sh);
  return v & 8191u;
}
// signed sum of n <= 8 registers (each < q), each taken +-(1..8) times (64 at most in total), fully reduced.  One warp per product means every
// instruction waits for its predecessor's latency, so the sum is built for instruction-level parallelism, not instruction count:
//   * carry-save accumulation: 12 independent 64-bit column sums, one multiply-add per limb (a subtracted limb enters as
//     x ^ 0xffffffff; the w_neg * 0xffffffff this adds to every column is taken off once at the end), starting from 64 q so
//     that the total is positive; the next source is fetched from shared memory while the current one is accumulated;
//   * quotient estimate k <= floor(total / q) <= k + 2 from the two top columns (a multiplication by a precomputed reciprocal of
//     q's top bits), k q taken off column-wise, ONE carry propagation, two conditional subtractions of q.
// tools/gen_pairing_prog.py runs the same integer procedure on real data and asserts 0 <= result < q.
HD_INLINE Fq381 pairing_linc(const Fq381* R, const PairingOp& op) {
  const uint32_t n = (op.w[0] >> 12) & 15u;
  unsigned long long acc[12];
  for (int i = 0; i < 12; i++) acc[i] = PairingProg::qoff(i);
  uint32_t w0 = (op.w[0] >> 16) | (op.w[1] << 16), w1 = (op.w[1] >> 16) | (op.w[2] << 16), w2 = (op.w[2] >> 16) | (op.w[3] << 16), w3 = op.w[3] >> 16;
  uint32_t wneg = 0;
  Fq381 x;
  copy_words16(&x, &R[w0 & 511u]);                  // n >= 1
#pragma unroll 1
  for (uint32_t j = 0; j < n; j++) {
    const uint32_t f = w0 & 8191u, m = 0u - ((f >> 9) & 1u), mult = (f >> 10) + 1u;
    w0 = (w0 >> 13) | (w1 << 19); w1 = (w1 >> 13) | (w2 << 19); w2 = (w2 >> 13) | (w3 << 19); w3 >>= 13;
    Fq381 nx;
    copy_words16(&nx, &R[w0 & 511u]);               // the next source (register 0 past the end: harmless)
#pragma unroll
    for (int i = 0; i < 12; i++) acc[i] += (unsigned long long)(x.v[i] ^ m) * mult;
    wneg += (m & 1u) * mult;
    x = nx;
  }
  // total / 2^352 lies in [T, T + 2): columns 11 and 12 (the latter only holds the top limb of 64 q) and the carry of column 10
  const long long fix = (long long)wneg * 0xffffffffll;
  const long long T = ((long long)acc[11] - fix) + (((long long)acc[10] - fix) >> 32) + ((long long)PairingProg::QOFF_TOP << 32);
  const uint32_t k = (uint32_t)((unsigned long long)((T >> 10) * (long long)PairingProg::Q_RECIP) >> 44);
  uint32_t r[12];
  long long c = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const long long v = (long long)acc[i] - fix - (long long)((unsigned long long)k * BlsFq::mod(i)) + c;
    r[i] = (uint32_t)v; c = v >> 32;
  }                                                  // c + QOFF_TOP = 0 by now: the total is below 3 q < 2^384
  cond_sub_p<BlsFq>(r, 0u);
  cond_sub_p<BlsFq>(r, 0u);
  Fq381 res;
  for (int i = 0; i < 12; i++) res.v[i] = r[i];
  return res;
}
HD_INLINE void pairing_lane_op(Fq381* R, const PairingOp& op) {
  const uint32_t kind = op.w[0] & 7u;
  if (kind == 0u) return;
  const uint32_t d = (op.w[0] >> 3) & 511u;
  Fq381 r;
  if (kind == 2u) r = pairing_linc(R, op);
  else {
    Fq381 x;
    copy_words16(&x, &R[pairing_op_field(op, 0) & 511u]);
    if (kind == 1u) { Fq381 y; copy_words16(&y, &R[pairing_op_field(op, 1) & 511u]); r = x * y; }
    else if (kind == 3u) r = fq381_half(x);
    else r = fq381_inv_fast(x);
  }
  copy_words16(&R[d], &r);
}
// the operation of `lane` in `step` (a step stores only its real operations; the other lanes idle)
HD_INLINE PairingOp pairing_op_load(uint32_t step, unsigned lane) {
  PairingOp op;
  const uint32_t first = PAIRING_STEP_OFF[step], n = PAIRING_STEP_OFF[step + 1] - first;
  if (lane < n) copy_words16(&op, reinterpret_cast<const PairingOp*>(PAIRING_OPS) + first + lane);
  else { op.w[0] = 0; op.w[1] = 0; op.w[2] = 0; op.w[3] = 0; }
  return op;
}

// the whole macro program.  Device: called by all threads of the block (lane = thread index); host (tests): one call plays all lanes.
HD_INLINE void pairing_prog_run(Fq381* R, unsigned lane) {
#if !defined(__CUDA_ARCH__)
  (void)lane;
#endif
  for (int m = 0; m < PairingProg::NMACRO; m++) {
    const uint32_t seg = PAIRING_MACRO[m];
    const uint32_t s0 = PAIRING_SEG_OFF[seg], s1 = PAIRING_SEG_OFF[seg + 1];
#if defined(__CUDA_ARCH__)
    PairingOp op = pairing_op_load(s0, lane);
#pragma unroll 1
    for (uint32_t s = s0; s < s1; s++) {
      PairingOp nxt = op;
      if (s + 1 < s1) nxt = pairing_op_load(s + 1, lane);      // in flight while this step computes
      pairing_lane_op(R, op);
      if (PairingProg::LANES == 32) __syncwarp(); else __syncthreads();
      op = nxt;
    }
#else
    for (uint32_t s = s0; s < s1; s++)
      for (int l = 0; l < PairingProg::LANES; l++) pairing_lane_op(R, pairing_op_load(s, (unsigned)l));
#endif
  }
}

// registers every program expects: 0, 1 and the Frobenius coefficients (the order of tools/gen_pairing_prog.py CONST_F2)
HD_INLINE void pairing_prog_load_consts(Fq381* R) {
  R[PairingProg::R_ZERO] = Fq381::zero();
  R[PairingProg::R_ONE] = Fq381::one();
  const Fq2 c[6] = {PAIRING_F2_CONST(G6_1_1), PAIRING_F2_CONST(G6_2_1), PAIRING_F2_CONST(G12_1),
                    PAIRING_F2_CONST(G6_1_2), PAIRING_F2_CONST(G6_2_2), PAIRING_F2_CONST(G12_2)};
  for (int k = 0; k < 6; k++) { R[PairingProg::R_CONST0 + 2 * k] = c[k].c0; R[PairingProg::R_CONST0 + 2 * k + 1] = c[k].c1; }
}
// pair k of the product: ABI bytes -> registers.  returns 0 ok, 1 the pair holds an identity (contributes 1), 2 malformed
HD_INLINE int pairing_prog_load_pair(Fq381* R, int k, const uint8_t* g1, const uint8_t* g2, bool negate) {
  G1AffPt p; G2Aff q;
  bool ok = g1_load_bytes(p, g1 + 96 * k) & g2_load(q, g2 + 192 * k);
  ok = ok && g1_on_curve_pt(p) && g2_on_curve(q);
  if (!ok) return 2;
  if (negate) p.y = neg(p.y);
  R[PairingProg::R_P0 + 2 * k] = p.x; R[PairingProg::R_P0 + 2 * k + 1] = p.y;
  R[PairingProg::R_Q0 + 4 * k] = q.x.c0; R[PairingProg::R_Q0 + 4 * k + 1] = q.x.c1;
  R[PairingProg::R_Q0 + 4 * k + 2] = q.y.c0; R[PairingProg::R_Q0 + 4 * k + 3] = q.y.c1;
  return (p.inf | q.inf) ? 1 : 0;
}
HD_INLINE Fq12 pairing_prog_result(const Fq381* R) {
  const Fq381* o = R + PairingProg::R_OUT;
  return Fq12{Fq6{Fq2{o[0], o[1]}, Fq2{o[2], o[3]}, Fq2{o[4], o[5]}}, Fq6{Fq2{o[6], o[7]}, Fq2{o[8], o[9]}, Fq2{o[10], o[11]}}};
}

#if !defined(__CUDA_ARCH__)
// host statement of the kernel below (tests/host_emul): the same verdicts as pairing_product_check_bytes for two pairs
HD_INLINE int pairing_product_check_lanes_host(const uint8_t* g1, const uint8_t* g2, unsigned negate, Fq12* value_out) {
  static Fq381 R[PairingProg::NREG];
  pairing_prog_load_consts(R);
  int st = 0;
  for (int k = 0; k < PairingProg::NPAIRS; k++) st |= pairing_prog_load_pair(R, k, g1, g2, (negate >> k) & 1u);
  if (st & 2) return 2;
  if (st & 1) return pairing_product_check_bytes(PairingProg::NPAIRS, g1, g2, negate, value_out);
  pairing_prog_run(R, 0);
  const Fq12 e = pairing_prog_result(R);
  if (value_out) *value_out = e;
  return f12_is_one(e) ? 1 : 0;
}
#endif

#ifdef __CUDACC__
// one block of PairingProg::LANES threads per product of PairingProg::NPAIRS pairings; same outputs as k_pairing_products.  bad_count (may be null): a non-zero
// counter of malformed inputs found by an earlier kernel of the call turns every verdict into 2 (the KZG check).
__global__ void __launch_bounds__(PairingProg::LANES) k_pairing_products_lanes(uint32_t n, const uint8_t* g1, const uint8_t* g2, const uint32_t* negate,
                                                               const uint32_t* bad_count, uint8_t* out_verdict, uint8_t* out_gt) {
  __shared__ Fq381 R[PairingProg::NREG];
  __shared__ int sh_state;
  const unsigned lane = threadIdx.x;

  constexpr int NP = PairingProg::NPAIRS;
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const uint8_t* p1 = g1 + (size_t)96 * NP * i;
    const uint8_t* p2 = g2 + (size_t)192 * NP * i;
    const unsigned neg_mask = negate ? negate[i] : 0u;
    if (lane == 0) sh_state = (bad_count && *bad_count) ? 4 : 0;
    __syncthreads();
    if ((int)lane < NP) { const int st = pairing_prog_load_pair(R, (int)lane, p1, p2, (neg_mask >> lane) & 1u); if (st) atomicOr(&sh_state, st); }
    if (lane == NP) pairing_prog_load_consts(R);
    __syncthreads();
    const int st = sh_state;
    if (st & 6) {
      if (lane == 0) { out_verdict[i] = 2; if (out_gt) f12_store(out_gt + (size_t)576 * i, f12_one()); }
    } else if (st & 1) {                               // an identity in the product: the one-thread form skips that pair
      if (lane == 0) {
        Fq12 e = f12_one();
        out_verdict[i] = (uint8_t)pairing_product_check_bytes(NP, p1, p2, neg_mask, &e);
        if (out_gt) f12_store(out_gt + (size_t)576 * i, e);
      }
    } else {
      pairing_prog_run(R, lane);
      if (lane == 0) {
        const Fq12 e = pairing_prog_result(R);
        out_verdict[i] = f12_is_one(e) ? 1 : 0;
        if (out_gt) f12_store(out_gt + (size_t)576 * i, e);
      }
    }
    __syncthreads();
  }
}
#endif

}  // namespace vrfs
