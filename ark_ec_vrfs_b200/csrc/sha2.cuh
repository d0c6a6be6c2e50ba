// K5 of SURVEY.md 2.5: SHA-512 / SHA-256 / HMAC-SHA-256 as per-thread streaming hashers (FIPS 180-4,
// RFC 2104) - the device replacement for the `sha2` / `hmac` crates behind `Suite::Hasher`
// (/root/reference/src/lib.rs:13-17).  Message blocks are kept as big-endian words so the compression
// function reads them directly; bytes are inserted with shifts (no byte-addressed local memory).
#pragma once
#include "arith.cuh"
#include "gen/sha2_consts.cuh"

namespace vrfs {

HD_INLINE uint64_t ror64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
HD_INLINE uint32_t ror32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

struct Sha512 {
  static constexpr int DIGEST = 64, BLOCK = 128;
  uint64_t h[8], w[16];
  uint32_t fill;   // bytes in the current block
  uint32_t total;  // message bytes so far (messages here are far below 2^32 bytes)

  HD_INLINE void init() {
    for (int i = 0; i < 8; i++) h[i] = SHA512_H0[i];
    for (int i = 0; i < 16; i++) w[i] = 0;
    fill = 0; total = 0;
  }
  HD_NOINLINE_M void compress() {
    uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    uint64_t x[16];
    for (int i = 0; i < 16; i++) x[i] = w[i];
#pragma unroll 1
    for (int r = 0; r < 80; r += 16) {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        if (r) {
          uint64_t w15 = x[(i + 1) & 15], w2 = x[(i + 14) & 15];
          uint64_t s0 = ror64(w15, 1) ^ ror64(w15, 8) ^ (w15 >> 7);
          uint64_t s1 = ror64(w2, 19) ^ ror64(w2, 61) ^ (w2 >> 6);
          x[i] += s0 + x[(i + 9) & 15] + s1;
        }
        uint64_t S1 = ror64(e, 14) ^ ror64(e, 18) ^ ror64(e, 41);
        uint64_t t1 = hh + S1 + ((e & f) ^ (~e & g)) + SHA512_K[r + i] + x[i];
        uint64_t S0 = ror64(a, 28) ^ ror64(a, 34) ^ ror64(a, 39);
        uint64_t t2 = S0 + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
      }
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    for (int i = 0; i < 16; i++) w[i] = 0;
    fill = 0;
  }
  HD_INLINE void put(uint8_t byte) {
    w[fill >> 3] |= (uint64_t)byte << (56 - 8 * (fill & 7));
    fill++; total++;
    if (fill == BLOCK) compress();
  }
  HD_INLINE void update(const uint8_t* p, uint32_t n) { for (uint32_t i = 0; i < n; i++) put(p[i]); }
  // out: 64 bytes
  HD_INLINE void final(uint8_t* out) {
    uint32_t bits = total * 8u;
    put(0x80);
    total--;
    if (fill > 112) compress();
    w[15] = bits;
    compress();
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(h[i] >> (56 - 8 * j));
  }
};

struct Sha256 {
  static constexpr int DIGEST = 32, BLOCK = 64;
  uint32_t h[8], w[16];
  uint32_t fill, total;

  HD_INLINE void init() {
    for (int i = 0; i < 8; i++) h[i] = SHA256_H0[i];
    for (int i = 0; i < 16; i++) w[i] = 0;
    fill = 0; total = 0;
  }
  HD_NOINLINE_M void compress() {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    uint32_t x[16];
    for (int i = 0; i < 16; i++) x[i] = w[i];
#pragma unroll 1
    for (int r = 0; r < 64; r += 16) {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        if (r) {
          uint32_t w15 = x[(i + 1) & 15], w2 = x[(i + 14) & 15];
          uint32_t s0 = ror32(w15, 7) ^ ror32(w15, 18) ^ (w15 >> 3);
          uint32_t s1 = ror32(w2, 17) ^ ror32(w2, 19) ^ (w2 >> 10);
          x[i] += s0 + x[(i + 9) & 15] + s1;
        }
        uint32_t S1 = ror32(e, 6) ^ ror32(e, 11) ^ ror32(e, 25);
        uint32_t t1 = hh + S1 + ((e & f) ^ (~e & g)) + SHA256_K[r + i] + x[i];
        uint32_t S0 = ror32(a, 2) ^ ror32(a, 13) ^ ror32(a, 22);
        uint32_t t2 = S0 + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
      }
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    for (int i = 0; i < 16; i++) w[i] = 0;
    fill = 0;
  }
  HD_INLINE void put(uint8_t byte) {
    w[fill >> 2] |= (uint32_t)byte << (24 - 8 * (fill & 3));
    fill++; total++;
    if (fill == BLOCK) compress();
  }
  HD_INLINE void update(const uint8_t* p, uint32_t n) { for (uint32_t i = 0; i < n; i++) put(p[i]); }
  HD_INLINE void final(uint8_t* out) {
    uint32_t bits = total * 8u;
    put(0x80);
    total--;
    if (fill > 56) compress();
    w[15] = bits;
    compress();
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) out[4 * i + j] = (uint8_t)(h[i] >> (24 - 8 * j));
  }
};

// HMAC-SHA-256 with a key of at most one block (RFC 6979 uses 32-byte keys)
struct HmacSha256 {
  Sha256 in, out;
  HD_INLINE void init(const uint8_t* key, uint32_t klen) {
    in.init(); out.init();
    for (uint32_t i = 0; i < 64; i++) { uint8_t k = i < klen ? key[i] : 0; in.put(k ^ 0x36); out.put(k ^ 0x5c); }
  }
  HD_INLINE void update(const uint8_t* p, uint32_t n) { in.update(p, n); }
  HD_INLINE void final(uint8_t* mac) {
    uint8_t t[32];
    in.final(t); out.update(t, 32); out.final(mac);
  }
};

}  // namespace vrfs
