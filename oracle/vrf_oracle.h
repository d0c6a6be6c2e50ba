/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, 64-bit limbs, unsigned __int128) of the algorithm behind the names the
 * reference re-exports at /root/reference/src/lib.rs:13-17 (`ark_vrf::{ietf, pedersen, ring, utils,
 * codec, suites, Suite, Secret, Public, Input, Output}`).  The implementation those names point to is
 * the crates.io dependency `ark-vrf = 0.1.0` (/root/reference/Cargo.toml:12) on top of arkworks 0.5
 * (ark-ff / ark-ec / ark-serialize), sha2, hmac and w3f ring-proof; NONE of these is mounted and no
 * Cargo.lock pins their versions (/root/reference/.gitignore:8).  The published algorithm restated
 * here is SURVEY.md Appendix A (RFC 9381 5.1-5.4, RFC 9380 5.3.1 + 6.8.2, RFC 8032-style nonce,
 * RFC 6979 3.2, plus the arkworks deviations pinned by reproduction).
 *
 * It deliberately follows the REFERENCE's algorithms, not the GPU engine's: Montgomery arithmetic on
 * 64-bit limbs, bit-serial double-and-add `mul_bigint`, one inversion per `into_affine`, ark-ec's
 * window-size rule for the Pippenger MSM.  That makes it (a) an independent checker of the CUDA
 * engine and (b) the "port" CPU baseline timed by bench.py.
 *
 * Parity status: pinned by tests/golden/bandersnatch_upstream.json (upstream Bandersnatch IETF x3,
 * Pedersen x1) and tests/golden/p256_rfc9381.json (RFC 9381 Appendix B Examples 10-11).
 * PARITY UNPINNED for: the Ed25519 suite, `ad` != "" placement, and the ring commitment layout
 * (no upstream vector available offline).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  Nothing under ark_ec_vrfs_b200/ links or calls it.
 *
 * Data formats are those of include/vrfs_b200.h (the product's C ABI) so the same buffers can be fed
 * to both: scalars 32 B little-endian canonical; points affine x||y, 32 B little-endian each (SW
 * identity = 64 zero bytes); BLS12-381 G1 points x||y 48 B little-endian each.
 */
#ifndef VRF_ORACLE_H
#define VRF_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_BANDERSNATCH = 0, ORACLE_ED25519 = 1, ORACLE_P256 = 2 };

int oracle_hash_len(int suite);      /* 64 (SHA-512 suites) or 32 */
int oracle_point_enc_len(int suite); /* 32 (arkworks codec) or 33 (SEC1) */
int oracle_challenge_len(int suite);

/* hashing primitives exposed for known-answer tests */
void oracle_sha512(const uint8_t *msg, size_t len, uint8_t out[64]);
void oracle_sha256(const uint8_t *msg, size_t len, uint8_t out[32]);
void oracle_hmac_sha256(const uint8_t *key, size_t klen, const uint8_t *msg, size_t len, uint8_t out[32]);

/* Secret::from_seed + public key.  seeds concatenated, seed_off[n+1]. */
void oracle_secret_from_seed_batch(int suite, size_t n, const uint8_t *seeds, const uint64_t *seed_off,
                                   uint8_t *out_sk, uint8_t *out_pk, int nthreads);
/* codec::point_encode / point_decode.  out_ok[i] = 1 on success. */
void oracle_point_encode_batch(int suite, size_t n, const uint8_t *pts, uint8_t *out_enc, int nthreads);
void oracle_point_decode_batch(int suite, size_t n, const uint8_t *enc, uint8_t *out_pts, uint8_t *out_ok, int nthreads);
/* Suite::data_to_point (Input::new).  out_ok[i] = 0 if no point was found. */
void oracle_data_to_point_batch(int suite, size_t n, const uint8_t *data, const uint64_t *data_off,
                                uint8_t *out_pts, uint8_t *out_ok, int nthreads);
/* Secret::output: O = sk * I */
void oracle_output_batch(int suite, size_t n, const uint8_t *sk, const uint8_t *input, uint8_t *out_output, int nthreads);
/* Output::hash / Suite::point_to_hash */
void oracle_point_to_hash_batch(int suite, size_t n, const uint8_t *pts, uint8_t *out_hash, int nthreads);
/* Suite::nonce */
void oracle_nonce_batch(int suite, size_t n, const uint8_t *sk, const uint8_t *input, uint8_t *out_k, int nthreads);
/* ietf::Prover::prove / ietf::Verifier::verify.  ad may be NULL (all empty). */
void oracle_ietf_prove_batch(int suite, size_t n, const uint8_t *sk, const uint8_t *input, const uint8_t *output,
                             const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_c, uint8_t *out_s, int nthreads);
void oracle_ietf_verify_batch(int suite, size_t n, const uint8_t *pk, const uint8_t *input, const uint8_t *output,
                              const uint8_t *c, const uint8_t *s, const uint8_t *ad, const uint64_t *ad_off,
                              uint8_t *out_ok, int nthreads);
/* pedersen::Prover::prove / Verifier::verify.  proof = pk_com || r || ok (3 x 64 B affine) || s || sb (2 x 32 B). */
void oracle_pedersen_prove_batch(int suite, size_t n, const uint8_t *sk, const uint8_t *input, const uint8_t *output,
                                 const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_proof, uint8_t *out_blinding,
                                 int nthreads);
void oracle_pedersen_verify_batch(int suite, size_t n, const uint8_t *input, const uint8_t *output, const uint8_t *proof,
                                  const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, int nthreads);
/* wire formats: ark-serialize deserialisation rules (canonical, on curve, prime-order subgroup via
 * mul_bigint(r).is_zero()), proof bytes c || s, signature = point_encode(Output) || c || s */
void oracle_subgroup_check_batch(int suite, size_t n, const uint8_t *pts, uint8_t *out_ok, int nthreads);
void oracle_point_decode_checked_batch(int suite, size_t n, const uint8_t *enc, uint8_t *out_pts, uint8_t *out_ok, int nthreads);
int oracle_ietf_signature_len(int suite);
void oracle_ietf_sign_wire_batch(int suite, size_t n, const uint8_t *sk, const uint8_t *data, const uint64_t *data_off,
                                 const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_sig, uint8_t *out_ok, int nthreads);
void oracle_ietf_verify_wire_batch(int suite, size_t n, const uint8_t *pk_enc, const uint8_t *data, const uint64_t *data_off, const uint8_t *sig,
                                   const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, uint8_t *out_hash, int nthreads);
int oracle_pedersen_signature_len(int suite);
void oracle_pedersen_sign_wire_batch(int suite, size_t n, const uint8_t *sk, const uint8_t *data, const uint64_t *data_off,
                                     const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_sig, uint8_t *out_blinding, uint8_t *out_ok, int nthreads);
void oracle_pedersen_verify_wire_batch(int suite, size_t n, const uint8_t *data, const uint64_t *data_off, const uint8_t *sig,
                                       const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, int nthreads);
/* the same verifiers reporting which `Error` variant the reference's Result<(), Error> carries (lib.rs:13-17 `Error`):
 * out_status[i] = 0 Ok(()), 1 Error::VerificationFailure (well-formed values, the proof does not check), 2 Error::InvalidData
 * (a value no typed Public / Input / Output / Proof can hold: non-canonical, off the curve, outside the prime-order subgroup on
 * the wire entry points, the un-encodable short-Weierstrass identity, no input point found). */
void oracle_ietf_verify_status_batch(int suite, size_t n, const uint8_t *pk, const uint8_t *input, const uint8_t *output,
                                     const uint8_t *c, const uint8_t *s, const uint8_t *ad, const uint64_t *ad_off,
                                     uint8_t *out_ok, uint8_t *out_status, int nthreads);
void oracle_pedersen_verify_status_batch(int suite, size_t n, const uint8_t *input, const uint8_t *output, const uint8_t *proof,
                                         const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, uint8_t *out_status, int nthreads);
void oracle_ietf_verify_wire_status_batch(int suite, size_t n, const uint8_t *pk_enc, const uint8_t *data, const uint64_t *data_off, const uint8_t *sig,
                                          const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, uint8_t *out_hash, uint8_t *out_status, int nthreads);
void oracle_pedersen_verify_wire_status_batch(int suite, size_t n, const uint8_t *data, const uint64_t *data_off, const uint8_t *sig,
                                              const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, uint8_t *out_status, int nthreads);
/* ark-ec VariableBaseMSM::msm over BLS12-381 G1: n_columns scalar columns over one base vector.
 * bases n*96 B, scalars n_columns*n*32 B (column-major), out n_columns*96 B (identity = zeros). */
void oracle_msm_g1(size_t n, const uint8_t *bases, const uint8_t *scalars, int n_columns, uint8_t *out, int nthreads);
/* k * G1 generator for a batch of scalars (builds synthetic SRS-like base vectors for tests/bench). */
void oracle_g1_mul_gen_batch(size_t n, const uint8_t *scalars, uint8_t *out_pts, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
