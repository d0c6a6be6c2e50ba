/* TEST INFRASTRUCTURE ONLY - see vrf_oracle.h for scope, provenance and parity status.
 * Section references "A.n" are to SURVEY.md Appendix A; "lib.rs:13-17" is
 * /root/reference/src/lib.rs:13-17, the only place the reference names these items. */
#include "vrf_oracle.h"
#include "sha2_consts.h"
#include "curve_consts.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

/* ======================================================================================
 * SHA-512 / SHA-256 / HMAC-SHA-256  (sha2 + hmac crates; FIPS 180-4, RFC 2104)
 * ====================================================================================== */
#define ROR64(x, n) (((x) >> (n)) | ((x) << (64 - (n))))
#define ROR32(x, n) (((x) >> (n)) | ((x) << (32 - (n))))

typedef struct { u64 h[8]; uint8_t buf[128]; size_t fill; u64 total; } sha512_ctx;
typedef struct { uint32_t h[8]; uint8_t buf[64]; size_t fill; u64 total; } sha256_ctx;

static void sha512_block(u64 h[8], const uint8_t *b) {
    u64 w[80];
    for (int i = 0; i < 16; i++) {
        u64 v = 0;
        for (int j = 0; j < 8; j++) v = (v << 8) | b[8 * i + j];
        w[i] = v;
    }
    for (int i = 16; i < 80; i++) {
        u64 s0 = ROR64(w[i - 15], 1) ^ ROR64(w[i - 15], 8) ^ (w[i - 15] >> 7);
        u64 s1 = ROR64(w[i - 2], 19) ^ ROR64(w[i - 2], 61) ^ (w[i - 2] >> 6);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    u64 a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 80; i++) {
        u64 S1 = ROR64(e, 14) ^ ROR64(e, 18) ^ ROR64(e, 41);
        u64 t1 = hh + S1 + ((e & f) ^ (~e & g)) + SHA512_K[i] + w[i];
        u64 S0 = ROR64(a, 28) ^ ROR64(a, 34) ^ ROR64(a, 39);
        u64 t2 = S0 + ((a & bb) ^ (a & c) ^ (bb & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
    }
    h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
static void sha512_init(sha512_ctx *c) { memcpy(c->h, SHA512_H0, 64); c->fill = 0; c->total = 0; }
static void sha512_update(sha512_ctx *c, const void *data, size_t len) {
    const uint8_t *p = data;
    c->total += len;
    while (len) {
        size_t k = 128 - c->fill; if (k > len) k = len;
        memcpy(c->buf + c->fill, p, k); c->fill += k; p += k; len -= k;
        if (c->fill == 128) { sha512_block(c->h, c->buf); c->fill = 0; }
    }
}
static void sha512_final(sha512_ctx *c, uint8_t out[64]) {
    u64 bits = c->total * 8;
    uint8_t pad = 0x80; sha512_update(c, &pad, 1);
    uint8_t z = 0; while (c->fill != 112) sha512_update(c, &z, 1);
    uint8_t len[16] = {0};
    for (int i = 0; i < 8; i++) len[15 - i] = (uint8_t)(bits >> (8 * i));
    sha512_update(c, len, 16);
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(c->h[i] >> (56 - 8 * j));
}
static void sha256_block(uint32_t h[8], const uint8_t *b) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)b[4 * i] << 24) | ((uint32_t)b[4 * i + 1] << 16) | ((uint32_t)b[4 * i + 2] << 8) | b[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = ROR32(w[i - 15], 7) ^ ROR32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = ROR32(w[i - 2], 17) ^ ROR32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = ROR32(e, 6) ^ ROR32(e, 11) ^ ROR32(e, 25);
        uint32_t t1 = hh + S1 + ((e & f) ^ (~e & g)) + SHA256_K[i] + w[i];
        uint32_t S0 = ROR32(a, 2) ^ ROR32(a, 13) ^ ROR32(a, 22);
        uint32_t t2 = S0 + ((a & bb) ^ (a & c) ^ (bb & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
    }
    h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
static void sha256_init(sha256_ctx *c) { memcpy(c->h, SHA256_H0, 32); c->fill = 0; c->total = 0; }
static void sha256_update(sha256_ctx *c, const void *data, size_t len) {
    const uint8_t *p = data;
    c->total += len;
    while (len) {
        size_t k = 64 - c->fill; if (k > len) k = len;
        memcpy(c->buf + c->fill, p, k); c->fill += k; p += k; len -= k;
        if (c->fill == 64) { sha256_block(c->h, c->buf); c->fill = 0; }
    }
}
static void sha256_final(sha256_ctx *c, uint8_t out[32]) {
    u64 bits = c->total * 8;
    uint8_t pad = 0x80; sha256_update(c, &pad, 1);
    uint8_t z = 0; while (c->fill != 56) sha256_update(c, &z, 1);
    uint8_t len[8];
    for (int i = 0; i < 8; i++) len[7 - i] = (uint8_t)(bits >> (8 * i));
    sha256_update(c, len, 8);
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) out[4 * i + j] = (uint8_t)(c->h[i] >> (24 - 8 * j));
}
void oracle_sha512(const uint8_t *m, size_t len, uint8_t out[64]) { sha512_ctx c; sha512_init(&c); sha512_update(&c, m, len); sha512_final(&c, out); }
void oracle_sha256(const uint8_t *m, size_t len, uint8_t out[32]) { sha256_ctx c; sha256_init(&c); sha256_update(&c, m, len); sha256_final(&c, out); }

typedef struct { sha256_ctx in, out; } hmac256_ctx;
static void hmac256_init(hmac256_ctx *h, const uint8_t *key, size_t klen) {
    uint8_t k[64] = {0}, pad[64];
    if (klen > 64) oracle_sha256(key, klen, k); else memcpy(k, key, klen);
    for (int i = 0; i < 64; i++) pad[i] = k[i] ^ 0x36;
    sha256_init(&h->in); sha256_update(&h->in, pad, 64);
    for (int i = 0; i < 64; i++) pad[i] = k[i] ^ 0x5c;
    sha256_init(&h->out); sha256_update(&h->out, pad, 64);
}
static void hmac256_final(hmac256_ctx *h, uint8_t out[32]) {
    uint8_t t[32]; sha256_final(&h->in, t); sha256_update(&h->out, t, 32); sha256_final(&h->out, out);
}
void oracle_hmac_sha256(const uint8_t *key, size_t klen, const uint8_t *msg, size_t len, uint8_t out[32]) {
    hmac256_ctx h; hmac256_init(&h, key, klen); sha256_update(&h.in, msg, len); hmac256_final(&h, out);
}

/* suite hasher: an incremental wrapper over either digest */
typedef struct { int is512; sha512_ctx a; sha256_ctx b; } hasher;
static void h_init(hasher *h, int is512) { h->is512 = is512; if (is512) sha512_init(&h->a); else sha256_init(&h->b); }
static void h_update(hasher *h, const void *d, size_t n) { if (h->is512) sha512_update(&h->a, d, n); else sha256_update(&h->b, d, n); }
static void h_final(hasher *h, uint8_t *out) { if (h->is512) sha512_final(&h->a, out); else sha256_final(&h->b, out); }

/* ======================================================================================
 * Prime fields: Montgomery form, 64-bit limbs (ark-ff Fp<MontBackend<_, N>>, N = 4 or 6)
 * ====================================================================================== */
#define MAXL 6
typedef struct { u64 v[MAXL]; } fe;
typedef struct {
    int n;            /* limbs */
    fe p;             /* modulus */
    fe one, r2;       /* R mod p, R^2 mod p */
    u64 ninv;         /* -p^-1 mod 2^64 */
    fe pm1h;          /* (p-1)/2, canonical */
    fe pm2;           /* p-2, canonical */
    int s;            /* two-adicity of p-1 */
    fe t;             /* odd part of p-1, canonical */
    fe tp1h;          /* (t+1)/2, canonical */
    fe zt;            /* z^t for a quadratic non-residue z, Montgomery form */
} fctx;

static int raw_cmp(const fctx *F, const fe *a, const fe *b) {
    for (int i = F->n - 1; i >= 0; i--) { if (a->v[i] != b->v[i]) return a->v[i] < b->v[i] ? -1 : 1; }
    return 0;
}
static u64 raw_add(int n, fe *r, const fe *a, const fe *b) {
    u64 c = 0;
    for (int i = 0; i < n; i++) { u128 t = (u128)a->v[i] + b->v[i] + c; r->v[i] = (u64)t; c = (u64)(t >> 64); }
    return c;
}
static u64 raw_sub(int n, fe *r, const fe *a, const fe *b) {
    u64 br = 0;
    for (int i = 0; i < n; i++) { u128 t = (u128)a->v[i] - b->v[i] - br; r->v[i] = (u64)t; br = (u64)(t >> 64) & 1; }
    return br;
}
static int f_is_zero(const fctx *F, const fe *a) { u64 o = 0; for (int i = 0; i < F->n; i++) o |= a->v[i]; return o == 0; }
static int f_eq(const fctx *F, const fe *a, const fe *b) { return raw_cmp(F, a, b) == 0; }
static void f_zero(fe *a) { memset(a, 0, sizeof *a); }
static void f_add(const fctx *F, fe *r, const fe *a, const fe *b) {
    fe t; u64 c = raw_add(F->n, &t, a, b);
    if (c || raw_cmp(F, &t, &F->p) >= 0) raw_sub(F->n, &t, &t, &F->p);
    *r = t;
}
static void f_sub(const fctx *F, fe *r, const fe *a, const fe *b) {
    fe t; if (raw_sub(F->n, &t, a, b)) raw_add(F->n, &t, &t, &F->p);
    *r = t;
}
static void f_neg(const fctx *F, fe *r, const fe *a) { fe z; f_zero(&z); f_sub(F, r, &z, a); }
static void f_dbl(const fctx *F, fe *r, const fe *a) { f_add(F, r, a, a); }
/* CIOS Montgomery product (the loop bounds are compile-time constants for the two limb counts in use, so the compiler unrolls
 * them the way ark-ff's generated N-limb code is unrolled) */
static inline __attribute__((always_inline)) void f_mul_n(const fctx *F, fe *r, const fe *a, const fe *b, const int n) {
    u64 t[MAXL + 2] = {0};
    for (int i = 0; i < n; i++) {
        u64 c = 0;
        for (int j = 0; j < n; j++) { u128 x = (u128)a->v[j] * b->v[i] + t[j] + c; t[j] = (u64)x; c = (u64)(x >> 64); }
        u128 x = (u128)t[n] + c; t[n] = (u64)x; t[n + 1] = (u64)(x >> 64);
        u64 m = t[0] * F->ninv;
        x = (u128)m * F->p.v[0] + t[0]; c = (u64)(x >> 64);
        for (int j = 1; j < n; j++) { x = (u128)m * F->p.v[j] + t[j] + c; t[j - 1] = (u64)x; c = (u64)(x >> 64); }
        x = (u128)t[n] + c; t[n - 1] = (u64)x; t[n] = t[n + 1] + (u64)(x >> 64);
    }
    fe o; f_zero(&o); for (int i = 0; i < n; i++) o.v[i] = t[i];
    if (t[n] || raw_cmp(F, &o, &F->p) >= 0) raw_sub(n, &o, &o, &F->p);
    *r = o;
}
static void f_mul(const fctx *F, fe *r, const fe *a, const fe *b) {
    if (F->n == 4) f_mul_n(F, r, a, b, 4); else if (F->n == 6) f_mul_n(F, r, a, b, 6); else f_mul_n(F, r, a, b, F->n);
}
static void f_sqr(const fctx *F, fe *r, const fe *a) { f_mul(F, r, a, a); }
static void f_from_raw(const fctx *F, fe *r, const fe *a) { f_mul(F, r, a, &F->r2); }  /* a < 2^(64n) */
static void f_to_raw(const fctx *F, fe *r, const fe *a) { fe o; f_zero(&o); o.v[0] = 1; f_mul(F, r, a, &o); }
static void f_set_u64(const fctx *F, fe *r, u64 x) { fe t; f_zero(&t); t.v[0] = x; f_from_raw(F, r, &t); }
static void f_pow(const fctx *F, fe *r, const fe *a, const fe *e) {
    fe acc = F->one, base = *a;
    int top = F->n * 64 - 1;
    while (top >= 0 && !((e->v[top / 64] >> (top % 64)) & 1)) top--;
    for (int i = top; i >= 0; i--) { f_sqr(F, &acc, &acc); if ((e->v[i / 64] >> (i % 64)) & 1) f_mul(F, &acc, &acc, &base); }
    *r = acc;
}
static void f_inv(const fctx *F, fe *r, const fe *a) { f_pow(F, r, a, &F->pm2); }   /* 0 -> 0 */
/* 1 if square (incl. 0), 0 otherwise */
static int f_is_square(const fctx *F, const fe *a) {
    if (f_is_zero(F, a)) return 1;
    fe t; f_pow(F, &t, a, &F->pm1h); return f_eq(F, &t, &F->one);
}
/* Tonelli-Shanks (ark-ff SqrtPrecomputation::TonelliShanks); returns 0 if non-residue */
static int f_sqrt(const fctx *F, fe *r, const fe *a) {
    if (f_is_zero(F, a)) { f_zero(r); return 1; }
    if (!f_is_square(F, a)) return 0;
    fe c = F->zt, t, x; int m = F->s;
    f_pow(F, &t, a, &F->t); f_pow(F, &x, a, &F->tp1h);
    while (!f_eq(F, &t, &F->one)) {
        int i = 0; fe t2 = t;
        while (!f_eq(F, &t2, &F->one)) { f_sqr(F, &t2, &t2); i++; }
        fe b = c; for (int k = 0; k < m - i - 1; k++) f_sqr(F, &b, &b);
        m = i; f_sqr(F, &c, &b); f_mul(F, &t, &t, &c); f_mul(F, &x, &x, &b);
    }
    *r = x; return 1;
}
/* canonical value > (p-1)/2 ?  (arkworks TEFlags "x is negative") */
static int f_is_high(const fctx *F, const fe *a) { fe t; f_to_raw(F, &t, a); return raw_cmp(F, &t, &F->pm1h) > 0; }
static int f_is_odd(const fctx *F, const fe *a) { fe t; f_to_raw(F, &t, a); return (int)(t.v[0] & 1); }

/* bytes -> field, reducing mod p (ark-ff from_{le,be}_bytes_mod_order); any length */
static void f_from_bytes_mod(const fctx *F, fe *r, const uint8_t *b, size_t len, int big_endian) {
    fe acc; f_zero(&acc);
    fe f256; f_set_u64(F, &f256, 256);
    for (size_t i = 0; i < len; i++) {
        uint8_t byte = big_endian ? b[i] : b[len - 1 - i];
        fe d; f_set_u64(F, &d, byte);
        f_mul(F, &acc, &acc, &f256); f_add(F, &acc, &acc, &d);
    }
    *r = acc;
}
/* canonical little-endian bytes (8*n of them) */
static void f_to_le(const fctx *F, const fe *a, uint8_t *out) {
    fe t; f_to_raw(F, &t, a);
    for (int i = 0; i < F->n; i++) for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(t.v[i] >> (8 * j));
}
static void f_to_be32(const fctx *F, const fe *a, uint8_t *out) { uint8_t t[48]; f_to_le(F, a, t); for (int i = 0; i < 32; i++) out[i] = t[31 - i]; }
/* canonical little-endian bytes -> field; returns 0 if value >= p */
static int f_from_le_canonical(const fctx *F, fe *r, const uint8_t *b) {
    fe t; f_zero(&t);
    for (int i = 0; i < F->n; i++) for (int j = 0; j < 8; j++) t.v[i] |= (u64)b[8 * i + j] << (8 * j);
    if (raw_cmp(F, &t, &F->p) >= 0) return 0;
    f_from_raw(F, r, &t); return 1;
}

static void hex_to_fe(fe *r, const char *hex) {
    f_zero(r); size_t L = strlen(hex);
    for (size_t i = 0; i < L; i++) {
        char ch = hex[L - 1 - i]; u64 d = ch <= '9' ? (u64)(ch - '0') : (u64)((ch | 32) - 'a' + 10);
        r->v[i / 16] |= d << (4 * (i % 16));
    }
}
static void raw_shr1(int n, fe *a) { for (int i = 0; i < n; i++) a->v[i] = (a->v[i] >> 1) | (i + 1 < n ? a->v[i + 1] << 63 : 0); }

static void fctx_init(fctx *F, int n, const char *p_hex) {
    memset(F, 0, sizeof *F); F->n = n; hex_to_fe(&F->p, p_hex);
    u64 inv = 1; for (int i = 0; i < 6; i++) inv *= 2 - F->p.v[0] * inv;   /* p^-1 mod 2^64 */
    F->ninv = (u64)0 - inv;
    fe x; f_zero(&x); x.v[0] = 1;                                          /* 2^(64n) and 2^(128n) mod p by doubling */
    for (int i = 0; i < 128 * n; i++) {
        u64 c = raw_add(n, &x, &x, &x);
        if (c || raw_cmp(F, &x, &F->p) >= 0) raw_sub(n, &x, &x, &F->p);
        if (i == 64 * n - 1) F->one = x;
    }
    F->r2 = x;
    fe one; f_zero(&one); one.v[0] = 1; fe two = one; two.v[0] = 2;
    raw_sub(n, &F->pm2, &F->p, &two);
    raw_sub(n, &F->pm1h, &F->p, &one); F->t = F->pm1h; raw_shr1(n, &F->pm1h);
    F->s = 0; while (!(F->t.v[0] & 1)) { raw_shr1(n, &F->t); F->s++; }
    raw_add(n, &F->tp1h, &F->t, &one); raw_shr1(n, &F->tp1h);
    for (u64 z = 2;; z++) { fe zm; f_set_u64(F, &zm, z); if (!f_is_square(F, &zm)) { f_pow(F, &F->zt, &zm, &F->t); break; } }
}

/* ======================================================================================
 * Curves.  One affine type for both models; `inf` only meaningful for short Weierstrass.
 * TE:  a x^2 + y^2 = 1 + d x^2 y^2, extended coordinates (ark-ec twisted_edwards::Projective)
 * SW:  y^2 = x^3 + a x + b, Jacobian coordinates (ark-ec short_weierstrass::Projective)
 * ====================================================================================== */
typedef struct { fe x, y; int inf; } aff;
typedef struct { fe X, Y, Z, T; } proj;   /* TE: extended; SW: Jacobian (T unused) */
typedef struct {
    int is_te;
    const fctx *F;     /* base field */
    const fctx *Fr;    /* scalar field */
    fe a, d_or_b;      /* Montgomery form */
    aff G;
    int cof_log2;
} curve;

static void pt_identity(const curve *C, proj *P) {
    const fctx *F = C->F; f_zero(&P->X); f_zero(&P->T);
    if (C->is_te) { P->Y = F->one; P->Z = F->one; } else { P->Y = F->one; f_zero(&P->Z); }
}
static void pt_from_aff(const curve *C, proj *P, const aff *A) {
    const fctx *F = C->F;
    if (!C->is_te && A->inf) { pt_identity(C, P); return; }
    P->X = A->x; P->Y = A->y; P->Z = F->one;
    if (C->is_te) f_mul(F, &P->T, &A->x, &A->y); else f_zero(&P->T);
}
static int pt_is_identity(const curve *C, const proj *P) {
    const fctx *F = C->F;
    if (C->is_te) return f_is_zero(F, &P->X) && f_eq(F, &P->Y, &P->Z);
    return f_is_zero(F, &P->Z);
}
static void pt_double(const curve *C, proj *R, const proj *P) {
    const fctx *F = C->F;
    if (C->is_te) {            /* dbl-2008-hwcd */
        fe A, B, Cc, D, E, G, Fv, H, t;
        f_sqr(F, &A, &P->X); f_sqr(F, &B, &P->Y); f_sqr(F, &Cc, &P->Z); f_dbl(F, &Cc, &Cc);
        f_mul(F, &D, &C->a, &A);
        f_add(F, &t, &P->X, &P->Y); f_sqr(F, &E, &t); f_sub(F, &E, &E, &A); f_sub(F, &E, &E, &B);
        f_add(F, &G, &D, &B); f_sub(F, &Fv, &G, &Cc); f_sub(F, &H, &D, &B);
        f_mul(F, &R->X, &E, &Fv); f_mul(F, &R->Y, &G, &H); f_mul(F, &R->T, &E, &H); f_mul(F, &R->Z, &Fv, &G);
    } else {                   /* dbl-2007-bl, general a */
        if (f_is_zero(F, &P->Z) || f_is_zero(F, &P->Y)) { pt_identity(C, R); return; }
        fe XX, YY, YYYY, ZZ, S, M, t, X3, Y3, Z3;
        f_sqr(F, &XX, &P->X); f_sqr(F, &YY, &P->Y); f_sqr(F, &YYYY, &YY); f_sqr(F, &ZZ, &P->Z);
        f_add(F, &t, &P->X, &YY); f_sqr(F, &t, &t); f_sub(F, &t, &t, &XX); f_sub(F, &t, &t, &YYYY); f_dbl(F, &S, &t);
        f_dbl(F, &M, &XX); f_add(F, &M, &M, &XX); f_sqr(F, &t, &ZZ); f_mul(F, &t, &t, &C->a); f_add(F, &M, &M, &t);
        f_sqr(F, &X3, &M); f_sub(F, &X3, &X3, &S); f_sub(F, &X3, &X3, &S);
        f_add(F, &Z3, &P->Y, &P->Z); f_sqr(F, &Z3, &Z3); f_sub(F, &Z3, &Z3, &YY); f_sub(F, &Z3, &Z3, &ZZ);
        f_sub(F, &t, &S, &X3); f_mul(F, &Y3, &M, &t);
        f_dbl(F, &t, &YYYY); f_dbl(F, &t, &t); f_dbl(F, &t, &t); f_sub(F, &Y3, &Y3, &t);
        R->X = X3; R->Y = Y3; R->Z = Z3;
    }
}
static void pt_add(const curve *C, proj *R, const proj *P, const proj *Q) {
    const fctx *F = C->F;
    if (C->is_te) {            /* add-2008-hwcd (unified, complete for this d) */
        fe A, B, Cc, D, E, Fv, G, H, t, u;
        f_mul(F, &A, &P->X, &Q->X); f_mul(F, &B, &P->Y, &Q->Y);
        f_mul(F, &Cc, &P->T, &Q->T); f_mul(F, &Cc, &Cc, &C->d_or_b); f_mul(F, &D, &P->Z, &Q->Z);
        f_add(F, &t, &P->X, &P->Y); f_add(F, &u, &Q->X, &Q->Y); f_mul(F, &E, &t, &u); f_sub(F, &E, &E, &A); f_sub(F, &E, &E, &B);
        f_sub(F, &Fv, &D, &Cc); f_add(F, &G, &D, &Cc); f_mul(F, &t, &C->a, &A); f_sub(F, &H, &B, &t);
        f_mul(F, &R->X, &E, &Fv); f_mul(F, &R->Y, &G, &H); f_mul(F, &R->T, &E, &H); f_mul(F, &R->Z, &Fv, &G);
    } else {                   /* add-2007-bl with the exceptional cases handled */
        if (f_is_zero(F, &P->Z)) { *R = *Q; return; }
        if (f_is_zero(F, &Q->Z)) { *R = *P; return; }
        fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, r, V, t, X3, Y3, Z3;
        f_sqr(F, &Z1Z1, &P->Z); f_sqr(F, &Z2Z2, &Q->Z);
        f_mul(F, &U1, &P->X, &Z2Z2); f_mul(F, &U2, &Q->X, &Z1Z1);
        f_mul(F, &S1, &P->Y, &Q->Z); f_mul(F, &S1, &S1, &Z2Z2);
        f_mul(F, &S2, &Q->Y, &P->Z); f_mul(F, &S2, &S2, &Z1Z1);
        if (f_eq(F, &U1, &U2)) { if (f_eq(F, &S1, &S2)) pt_double(C, R, P); else pt_identity(C, R); return; }
        f_sub(F, &H, &U2, &U1); f_dbl(F, &I, &H); f_sqr(F, &I, &I); f_mul(F, &J, &H, &I);
        f_sub(F, &r, &S2, &S1); f_dbl(F, &r, &r); f_mul(F, &V, &U1, &I);
        f_sqr(F, &X3, &r); f_sub(F, &X3, &X3, &J); f_sub(F, &X3, &X3, &V); f_sub(F, &X3, &X3, &V);
        f_sub(F, &t, &V, &X3); f_mul(F, &Y3, &r, &t); f_mul(F, &t, &S1, &J); f_dbl(F, &t, &t); f_sub(F, &Y3, &Y3, &t);
        f_add(F, &Z3, &P->Z, &Q->Z); f_sqr(F, &Z3, &Z3); f_sub(F, &Z3, &Z3, &Z1Z1); f_sub(F, &Z3, &Z3, &Z2Z2); f_mul(F, &Z3, &Z3, &H);
        R->X = X3; R->Y = Y3; R->Z = Z3;
    }
}
static void pt_neg(const curve *C, proj *R, const proj *P) {
    const fctx *F = C->F; *R = *P;
    if (C->is_te) { f_neg(F, &R->X, &P->X); f_neg(F, &R->T, &P->T); } else f_neg(F, &R->Y, &P->Y);
}
/* into_affine: one inversion per point, like the reference */
static void pt_to_aff(const curve *C, aff *A, const proj *P) {
    const fctx *F = C->F; fe zi;
    memset(A, 0, sizeof *A);
    if (C->is_te) { f_inv(F, &zi, &P->Z); f_mul(F, &A->x, &P->X, &zi); f_mul(F, &A->y, &P->Y, &zi); return; }
    if (f_is_zero(F, &P->Z)) { A->inf = 1; return; }
    fe zi2; f_inv(F, &zi, &P->Z); f_sqr(F, &zi2, &zi); f_mul(F, &A->x, &P->X, &zi2); f_mul(F, &zi2, &zi2, &zi); f_mul(F, &A->y, &P->Y, &zi2);
}
/* ark-ec `mul_bigint`: MSB-first double-and-add over the canonical scalar (n64 limbs) */
static void pt_mul(const curve *C, proj *R, const u64 *k, int n64, const proj *P) {
    proj acc; pt_identity(C, &acc);
    int top = n64 * 64 - 1;
    while (top >= 0 && !((k[top / 64] >> (top % 64)) & 1)) top--;
    for (int i = top; i >= 0; i--) { pt_double(C, &acc, &acc); if ((k[i / 64] >> (i % 64)) & 1) pt_add(C, &acc, &acc, P); }
    *R = acc;
}
static void aff_mul(const curve *C, aff *R, const fe *k_raw, const aff *P) {
    proj p, r; pt_from_aff(C, &p, P); pt_mul(C, &r, k_raw->v, C->Fr->n, &p); pt_to_aff(C, R, &r);
}
static int aff_on_curve(const curve *C, const aff *A) {
    const fctx *F = C->F; fe l, r, xx, yy, t;
    if (!C->is_te && A->inf) return 1;
    f_sqr(F, &xx, &A->x); f_sqr(F, &yy, &A->y);
    if (C->is_te) { f_mul(F, &l, &C->a, &xx); f_add(F, &l, &l, &yy); f_mul(F, &t, &xx, &yy); f_mul(F, &t, &t, &C->d_or_b); f_add(F, &r, &F->one, &t); }
    else { l = yy; f_mul(F, &r, &xx, &A->x); f_mul(F, &t, &C->a, &A->x); f_add(F, &r, &r, &t); f_add(F, &r, &r, &C->d_or_b); }
    return f_eq(F, &l, &r);
}
static int aff_eq(const curve *C, const aff *A, const aff *B) {
    if (!C->is_te && (A->inf || B->inf)) return A->inf == B->inf;
    return f_eq(C->F, &A->x, &B->x) && f_eq(C->F, &A->y, &B->y);
}

/* ======================================================================================
 * Global parameter tables (hex literals generated into curve_consts.h)
 * ====================================================================================== */
static fctx F_BLSFR, F_BANDR, F_25519, F_EDL, F_P256, F_P256N, F_BLSFQ, F_JUBR, F_BN254R, F_BJJR;
static curve C_BAND, C_ED, C_P256, C_G1, C_BSW, C_JUB, C_BJJ;
typedef struct {
    const curve *C; const uint8_t *suite_id; size_t suite_id_len; int clen, is512, sec1, ell2, rfc6979; aff B;
} suite_t;
#define N_SUITES 6
static suite_t SUITES[N_SUITES];
/* encoded point length: 32 B for the arkworks twisted-Edwards form; 33 B for SEC1 and for the arkworks short-Weierstrass form */
static size_t pt_len(const suite_t *S) { return (S->sec1 || !S->C->is_te) ? 33 : 32; }
static fe ELL2_JK, ELL2_KSQI, ELL2_K, ELL2_Z;   /* J/K, 1/K^2, K, Z=5 (A.5) */
static pthread_once_t init_once = PTHREAD_ONCE_INIT;

static void set_aff_hex(const fctx *F, aff *A, const char *xh, const char *yh) {
    fe t; memset(A, 0, sizeof *A); hex_to_fe(&t, xh); f_from_raw(F, &A->x, &t); hex_to_fe(&t, yh); f_from_raw(F, &A->y, &t);
}
static void init_all(void) {
    fctx_init(&F_BLSFR, 4, HEX_BLS_FR); fctx_init(&F_BANDR, 4, HEX_BAND_R);
    fctx_init(&F_25519, 4, HEX_P25519); fctx_init(&F_EDL, 4, HEX_ED_L);
    fctx_init(&F_P256, 4, HEX_P256_P); fctx_init(&F_P256N, 4, HEX_P256_N);
    fctx_init(&F_BLSFQ, 6, HEX_BLS_FQ);
    fe t;
    memset(&C_BAND, 0, sizeof C_BAND); C_BAND.is_te = 1; C_BAND.F = &F_BLSFR; C_BAND.Fr = &F_BANDR; C_BAND.cof_log2 = 2;
    f_set_u64(&F_BLSFR, &t, 5); f_neg(&F_BLSFR, &C_BAND.a, &t);
    hex_to_fe(&t, HEX_BAND_D); f_from_raw(&F_BLSFR, &C_BAND.d_or_b, &t);
    set_aff_hex(&F_BLSFR, &C_BAND.G, HEX_BAND_GX, HEX_BAND_GY);

    memset(&C_ED, 0, sizeof C_ED); C_ED.is_te = 1; C_ED.F = &F_25519; C_ED.Fr = &F_EDL; C_ED.cof_log2 = 3;
    f_neg(&F_25519, &C_ED.a, &F_25519.one);
    hex_to_fe(&t, HEX_ED_D); f_from_raw(&F_25519, &C_ED.d_or_b, &t);
    set_aff_hex(&F_25519, &C_ED.G, HEX_ED_GX, HEX_ED_GY);

    memset(&C_P256, 0, sizeof C_P256); C_P256.F = &F_P256; C_P256.Fr = &F_P256N;
    f_set_u64(&F_P256, &t, 3); f_neg(&F_P256, &C_P256.a, &t);
    hex_to_fe(&t, HEX_P256_B); f_from_raw(&F_P256, &C_P256.d_or_b, &t);
    set_aff_hex(&F_P256, &C_P256.G, HEX_P256_GX, HEX_P256_GY);

    memset(&C_G1, 0, sizeof C_G1); C_G1.F = &F_BLSFQ; C_G1.Fr = &F_BLSFR;
    f_zero(&C_G1.a); f_set_u64(&F_BLSFQ, &C_G1.d_or_b, 4);
    set_aff_hex(&F_BLSFQ, &C_G1.G, HEX_G1_X, HEX_G1_Y);

    static const uint8_t id_band[] = "Bandersnatch_SHA-512_ELL2", id_ed[] = "Ed25519_SHA-512_TAI", id_p256[] = {0x01};
    memset(SUITES, 0, sizeof SUITES);
    SUITES[0].C = &C_BAND; SUITES[0].suite_id = id_band; SUITES[0].suite_id_len = 25; SUITES[0].clen = 32; SUITES[0].is512 = 1; SUITES[0].ell2 = 1;
    SUITES[1].C = &C_ED;   SUITES[1].suite_id = id_ed;   SUITES[1].suite_id_len = 19; SUITES[1].clen = 16; SUITES[1].is512 = 1;
    SUITES[2].C = &C_P256; SUITES[2].suite_id = id_p256; SUITES[2].suite_id_len = 1;  SUITES[2].clen = 16; SUITES[2].sec1 = 1; SUITES[2].rfc6979 = 1;
    set_aff_hex(&F_BLSFR, &SUITES[0].B, HEX_BAND_BX, HEX_BAND_BY);
    set_aff_hex(&F_25519, &SUITES[1].B, HEX_ED_BX, HEX_ED_BY);
    set_aff_hex(&F_P256, &SUITES[2].B, HEX_P256_BX, HEX_P256_BY);

    /* bandersnatch_sw, jubjub, baby-jubjub - constants checked in oracle/pyref.py; suite strings [RECALL];
     * blinding bases are placeholders (pyref._placeholder_blinding_base).  PARITY UNPINNED. */
    fctx_init(&F_JUBR, 4, HEX_JUB_R); fctx_init(&F_BN254R, 4, HEX_BN254_FR); fctx_init(&F_BJJR, 4, HEX_BJJ_R);
    memset(&C_BSW, 0, sizeof C_BSW); C_BSW.F = &F_BLSFR; C_BSW.Fr = &F_BANDR; C_BSW.cof_log2 = 2;
    hex_to_fe(&t, HEX_BSW_A); f_from_raw(&F_BLSFR, &C_BSW.a, &t); hex_to_fe(&t, HEX_BSW_B); f_from_raw(&F_BLSFR, &C_BSW.d_or_b, &t);
    set_aff_hex(&F_BLSFR, &C_BSW.G, HEX_BSW_GX, HEX_BSW_GY);
    memset(&C_JUB, 0, sizeof C_JUB); C_JUB.is_te = 1; C_JUB.F = &F_BLSFR; C_JUB.Fr = &F_JUBR; C_JUB.cof_log2 = 3;
    f_neg(&F_BLSFR, &C_JUB.a, &F_BLSFR.one); hex_to_fe(&t, HEX_JUB_D); f_from_raw(&F_BLSFR, &C_JUB.d_or_b, &t);
    set_aff_hex(&F_BLSFR, &C_JUB.G, HEX_JUB_GX, HEX_JUB_GY);
    memset(&C_BJJ, 0, sizeof C_BJJ); C_BJJ.is_te = 1; C_BJJ.F = &F_BN254R; C_BJJ.Fr = &F_BJJR; C_BJJ.cof_log2 = 3;
    C_BJJ.a = F_BN254R.one; hex_to_fe(&t, HEX_BJJ_D); f_from_raw(&F_BN254R, &C_BJJ.d_or_b, &t);
    set_aff_hex(&F_BN254R, &C_BJJ.G, HEX_BJJ_GX, HEX_BJJ_GY);
    static const uint8_t id_bsw[] = "Bandersnatch_SW_SHA-512_TAI", id_jub[] = "JubJub_SHA-512_TAI", id_bjj[] = "BabyJubJub_SHA-512_TAI";
    SUITES[3].C = &C_BSW; SUITES[3].suite_id = id_bsw; SUITES[3].suite_id_len = 27; SUITES[3].clen = 32; SUITES[3].is512 = 1;
    SUITES[4].C = &C_JUB; SUITES[4].suite_id = id_jub; SUITES[4].suite_id_len = 18; SUITES[4].clen = 32; SUITES[4].is512 = 1;
    SUITES[5].C = &C_BJJ; SUITES[5].suite_id = id_bjj; SUITES[5].suite_id_len = 22; SUITES[5].clen = 32; SUITES[5].is512 = 1;
    set_aff_hex(&F_BLSFR, &SUITES[3].B, HEX_BSW_BX, HEX_BSW_BY);
    set_aff_hex(&F_BLSFR, &SUITES[4].B, HEX_JUB_BX, HEX_JUB_BY);
    set_aff_hex(&F_BN254R, &SUITES[5].B, HEX_BJJ_BX, HEX_BJJ_BY);

    fe A_, B_, bi;
    hex_to_fe(&t, HEX_BAND_MONT_A); f_from_raw(&F_BLSFR, &A_, &t);
    hex_to_fe(&t, HEX_BAND_MONT_B); f_from_raw(&F_BLSFR, &B_, &t);
    f_inv(&F_BLSFR, &bi, &B_); f_mul(&F_BLSFR, &ELL2_JK, &A_, &bi); f_sqr(&F_BLSFR, &ELL2_KSQI, &bi);
    ELL2_K = B_; f_set_u64(&F_BLSFR, &ELL2_Z, 5);
}
static const suite_t *get_suite(int id) { pthread_once(&init_once, init_all); return &SUITES[id]; }
int oracle_hash_len(int suite) { return get_suite(suite)->is512 ? 64 : 32; }
int oracle_point_enc_len(int suite) { return (int)pt_len(get_suite(suite)); }
int oracle_challenge_len(int suite) { return get_suite(suite)->clen; }

/* ======================================================================================
 * ABI <-> internal conversions (formats of include/vrfs_b200.h)
 * ====================================================================================== */
static int load_point(const curve *C, aff *A, const uint8_t *b) {      /* x||y LE; returns 0 if not canonical */
    int nb = C->F->n * 8; memset(A, 0, sizeof *A);
    if (!C->is_te) { int z = 1; for (int i = 0; i < 2 * nb; i++) if (b[i]) { z = 0; break; } if (z) { A->inf = 1; return 1; } }
    return f_from_le_canonical(C->F, &A->x, b) && f_from_le_canonical(C->F, &A->y, b + nb);
}
static void store_point(const curve *C, const aff *A, uint8_t *b) {
    int nb = C->F->n * 8;
    if (!C->is_te && A->inf) { memset(b, 0, 2 * nb); return; }
    f_to_le(C->F, &A->x, b); f_to_le(C->F, &A->y, b + nb);
}
/* scalar: 32 B LE, reduced mod r on load (codec scalar_decode = from_le_bytes_mod_order); kept canonical (non-Montgomery) */
static void load_scalar(const curve *C, fe *k, const uint8_t *b) { fe m; f_from_bytes_mod(C->Fr, &m, b, 32, 0); f_to_raw(C->Fr, k, &m); }
static void store_scalar(fe *k, uint8_t *b) { for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) b[8 * i + j] = (uint8_t)(k->v[i] >> (8 * j)); }

/* ======================================================================================
 * codec  (ark_vrf::codec, lib.rs:13-17; A.2)
 * ====================================================================================== */
static size_t enc_point(const suite_t *S, const aff *A, uint8_t *out) {
    const fctx *F = S->C->F;
    if (S->sec1) { out[0] = (uint8_t)(2 + f_is_odd(F, &A->y)); f_to_be32(F, &A->x, out + 1); return 33; }
    if (!S->C->is_te) {   /* arkworks short-Weierstrass compressed form [RECALL]: x LE (32 B) + flag byte: bit 7 = y > (p-1)/2, bit 6 = infinity */
        if (A->inf) { memset(out, 0, 33); out[32] = 0x40; return 33; }
        f_to_le(F, &A->x, out); out[32] = f_is_high(F, &A->y) ? 0x80 : 0x00; return 33;
    }
    f_to_le(F, &A->y, out); if (f_is_high(F, &A->x)) out[31] |= 0x80; return 32;
}
static size_t enc_scalar(const suite_t *S, const fe *k_raw, uint8_t *out) {
    uint8_t t[32]; fe k = *k_raw; store_scalar(&k, t);
    if (S->sec1) for (int i = 0; i < 32; i++) out[i] = t[31 - i]; else memcpy(out, t, 32);
    return 32;
}
static int dec_point(const suite_t *S, aff *A, const uint8_t *in) {   /* on-curve, no subgroup check */
    const curve *C = S->C; const fctx *F = C->F; memset(A, 0, sizeof *A);
    if (S->sec1) {
        if (in[0] != 2 && in[0] != 3) return 0;
        uint8_t le[32]; for (int i = 0; i < 32; i++) le[i] = in[32 - i];
        if (!f_from_le_canonical(F, &A->x, le)) return 0;
        fe rhs, t; f_sqr(F, &rhs, &A->x); f_mul(F, &rhs, &rhs, &A->x); f_mul(F, &t, &C->a, &A->x); f_add(F, &rhs, &rhs, &t); f_add(F, &rhs, &rhs, &C->d_or_b);
        if (!f_sqrt(F, &A->y, &rhs)) return 0;
        if (f_is_odd(F, &A->y) != (in[0] & 1)) f_neg(F, &A->y, &A->y);
        return 1;
    }
    if (!C->is_te) {
        int flags = in[32] >> 6;
        if (flags == 3) return 0;
        if (!f_from_le_canonical(F, &A->x, in)) return 0;
        if (flags == 1) return 0;                   /* the identity: no typed Public / Input / Output holds it */
        fe rhs, t; f_sqr(F, &rhs, &A->x); f_mul(F, &rhs, &rhs, &A->x); f_mul(F, &t, &C->a, &A->x); f_add(F, &rhs, &rhs, &t); f_add(F, &rhs, &rhs, &C->d_or_b);
        if (!f_sqrt(F, &A->y, &rhs)) return 0;
        if (f_is_high(F, &A->y) != ((flags >> 1) & 1)) f_neg(F, &A->y, &A->y);
        return 1;
    }
    uint8_t b[32]; memcpy(b, in, 32); int sign = b[31] >> 7; b[31] &= 0x7f;
    if (!f_from_le_canonical(F, &A->y, b)) return 0;
    fe yy, num, den, x2; f_sqr(F, &yy, &A->y); f_sub(F, &num, &F->one, &yy);
    f_mul(F, &den, &C->d_or_b, &yy); f_sub(F, &den, &C->a, &den);
    if (f_is_zero(F, &den)) return 0;
    f_inv(F, &den, &den); f_mul(F, &x2, &num, &den);
    if (!f_sqrt(F, &A->x, &x2)) return 0;
    if (f_is_high(F, &A->x) != sign) f_neg(F, &A->x, &A->x);
    return 1;
}

/* ======================================================================================
 * hash-to-curve  (utils::hash_to_curve_ell2_rfc_9380 / hash_to_curve_tai_rfc_9381; A.5)
 * ====================================================================================== */
static void elligator2(const curve *C, aff *out, const fe *u) {
    const fctx *F = C->F; fe den, x1, x2, gx1, gx2, t, x, y; int sgn;
    f_sqr(F, &den, u); f_mul(F, &den, &den, &ELL2_Z); f_add(F, &den, &den, &F->one);
    if (f_is_zero(F, &den)) den = F->one;
    f_inv(F, &den, &den); f_mul(F, &x1, &ELL2_JK, &den); f_neg(F, &x1, &x1);
    /* g(x) = x^3 + (J/K) x^2 + x/K^2 */
    f_sqr(F, &t, &x1); f_mul(F, &gx1, &t, &x1); f_mul(F, &t, &t, &ELL2_JK); f_add(F, &gx1, &gx1, &t); f_mul(F, &t, &x1, &ELL2_KSQI); f_add(F, &gx1, &gx1, &t);
    f_neg(F, &x2, &x1); f_sub(F, &x2, &x2, &ELL2_JK);
    f_sqr(F, &t, &x2); f_mul(F, &gx2, &t, &x2); f_mul(F, &t, &t, &ELL2_JK); f_add(F, &gx2, &gx2, &t); f_mul(F, &t, &x2, &ELL2_KSQI); f_add(F, &gx2, &gx2, &t);
    if (f_is_square(F, &gx1)) { x = x1; f_sqrt(F, &y, &gx1); sgn = 1; } else { x = x2; f_sqrt(F, &y, &gx2); sgn = 0; }
    if (f_is_odd(F, &y) != sgn) f_neg(F, &y, &y);
    fe s, tt, tv1, tv2; f_mul(F, &s, &x, &ELL2_K); f_mul(F, &tt, &y, &ELL2_K);
    f_add(F, &tv1, &s, &F->one); f_mul(F, &tv2, &tv1, &tt);
    memset(out, 0, sizeof *out);
    if (f_is_zero(F, &tv2)) { f_zero(&out->x); out->y = F->one; return; }
    fe ti, vi; f_inv(F, &ti, &tt); f_mul(F, &out->x, &s, &ti);
    f_inv(F, &vi, &tv1); f_sub(F, &t, &s, &F->one); f_mul(F, &out->y, &t, &vi);
}
static int h2c_ell2(const suite_t *S, aff *out, const uint8_t *data, size_t len) {
    /* DST = "ECVRF_" || h2c_suite_id || SUITE_ID ; expand_message_xmd with ark-ff's 48-byte Z_pad */
    static const char h2c_id[] = "Bandersnatch_XMD:SHA-512_ELL2_RO_";
    uint8_t dstp[96]; size_t dl = 0;
    memcpy(dstp, "ECVRF_", 6); dl = 6; memcpy(dstp + dl, h2c_id, 33); dl += 33; memcpy(dstp + dl, S->suite_id, S->suite_id_len); dl += S->suite_id_len;
    dstp[dl] = (uint8_t)dl; dl++;
    uint8_t zpad[48] = {0}, lib[3] = {0x00, 0x60, 0x00}, b0[64], b1[64], b2[64], x[64], one = 1, two = 2;
    sha512_ctx c; sha512_init(&c); sha512_update(&c, zpad, 48); sha512_update(&c, data, len); sha512_update(&c, lib, 3); sha512_update(&c, dstp, dl); sha512_final(&c, b0);
    sha512_init(&c); sha512_update(&c, b0, 64); sha512_update(&c, &one, 1); sha512_update(&c, dstp, dl); sha512_final(&c, b1);
    for (int i = 0; i < 64; i++) x[i] = b0[i] ^ b1[i];
    sha512_init(&c); sha512_update(&c, x, 64); sha512_update(&c, &two, 1); sha512_update(&c, dstp, dl); sha512_final(&c, b2);
    uint8_t uni[128]; memcpy(uni, b1, 64); memcpy(uni + 64, b2, 64);
    const curve *C = S->C; fe u0, u1; aff q0, q1;
    f_from_bytes_mod(C->F, &u0, uni, 48, 1); f_from_bytes_mod(C->F, &u1, uni + 48, 48, 1);
    elligator2(C, &q0, &u0); elligator2(C, &q1, &u1);
    proj p0, p1; pt_from_aff(C, &p0, &q0); pt_from_aff(C, &p1, &q1); pt_add(C, &p0, &p0, &p1);
    for (int i = 0; i < C->cof_log2; i++) pt_double(C, &p0, &p0);      /* clear_cofactor */
    pt_to_aff(C, out, &p0);
    return 1;
}
static int h2c_tai(const suite_t *S, aff *out, const uint8_t *data, size_t len) {
    const curve *C = S->C; uint8_t hs[65], front = 0x01, back = 0x00;
    for (int ctr = 0; ctr < 256; ctr++) {
        uint8_t cb = (uint8_t)ctr; hasher h; h_init(&h, S->is512);
        h_update(&h, S->suite_id, S->suite_id_len); h_update(&h, &front, 1); h_update(&h, data, len); h_update(&h, &cb, 1); h_update(&h, &back, 1);
        h_final(&h, hs + 1);
        aff P; int ok;
        if (S->sec1) { hs[0] = 0x02; ok = dec_point(S, &P, hs); } else ok = dec_point(S, &P, hs + 1);
        if (!ok) continue;
        proj p; pt_from_aff(C, &p, &P);
        for (int i = 0; i < C->cof_log2; i++) pt_double(C, &p, &p);
        if (pt_is_identity(C, &p)) continue;
        pt_to_aff(C, out, &p); return 1;
    }
    return 0;
}
static int data_to_point(const suite_t *S, aff *out, const uint8_t *data, size_t len) {
    return S->ell2 ? h2c_ell2(S, out, data, len) : h2c_tai(S, out, data, len);
}

/* ======================================================================================
 * nonce / challenge / point_to_hash / blinding  (utils::*, lib.rs:13-17; A.6-A.8, A.10)
 * ====================================================================================== */
static void nonce_rfc8032(const suite_t *S, fe *k, const fe *sk_raw, const aff *I) {
    uint8_t e[32], h1[64], h2[64], ip[33]; enc_scalar(S, sk_raw, e);
    hasher h; h_init(&h, S->is512); h_update(&h, e, 32); h_final(&h, h1);
    size_t il = enc_point(S, I, ip);
    h_init(&h, S->is512); h_update(&h, h1 + 32, 32); h_update(&h, ip, il); h_final(&h, h2);
    fe m; f_from_bytes_mod(S->C->Fr, &m, h2, 64, 0); f_to_raw(S->C->Fr, k, &m);
}
static void nonce_rfc6979(const suite_t *S, fe *k, const fe *sk_raw, const aff *I) {
    const fctx *Fr = S->C->Fr; uint8_t ip[33], h1[32], xb[32], hb[32], V[32], K[32], sep;
    size_t il = enc_point(S, I, ip); oracle_sha256(ip, il, h1);
    fe hm, hr; f_from_bytes_mod(Fr, &hm, h1, 32, 1); f_to_raw(Fr, &hr, &hm);       /* bits2octets */
    uint8_t t[32]; store_scalar(&hr, t); for (int i = 0; i < 32; i++) hb[i] = t[31 - i];
    fe s = *sk_raw; store_scalar(&s, t); for (int i = 0; i < 32; i++) xb[i] = t[31 - i];
    memset(V, 1, 32); memset(K, 0, 32);
    hmac256_ctx h;
    for (sep = 0; sep < 2; sep++) {
        hmac256_init(&h, K, 32); sha256_update(&h.in, V, 32); sha256_update(&h.in, &sep, 1); sha256_update(&h.in, xb, 32); sha256_update(&h.in, hb, 32); hmac256_final(&h, K);
        hmac256_init(&h, K, 32); sha256_update(&h.in, V, 32); hmac256_final(&h, V);
    }
    for (;;) {
        hmac256_init(&h, K, 32); sha256_update(&h.in, V, 32); hmac256_final(&h, V);
        fe cand; f_zero(&cand);
        for (int i = 0; i < 32; i++) cand.v[(31 - i) / 8] |= (u64)V[i] << (8 * ((31 - i) % 8));
        if (!f_is_zero(Fr, &cand) && raw_cmp(Fr, &cand, &Fr->p) < 0) { *k = cand; return; }
        sep = 0;
        hmac256_init(&h, K, 32); sha256_update(&h.in, V, 32); sha256_update(&h.in, &sep, 1); hmac256_final(&h, K);
        hmac256_init(&h, K, 32); sha256_update(&h.in, V, 32); hmac256_final(&h, V);
    }
}
static void suite_nonce(const suite_t *S, fe *k, const fe *sk_raw, const aff *I) {
    if (S->rfc6979) nonce_rfc6979(S, k, sk_raw, I); else nonce_rfc8032(S, k, sk_raw, I);
}
static void suite_challenge(const suite_t *S, fe *c_raw, const aff *const pts[5], const uint8_t *ad, size_t adlen) {
    uint8_t two = 0x02, zero = 0x00, e[33], d[64]; hasher h; h_init(&h, S->is512);
    h_update(&h, S->suite_id, S->suite_id_len); h_update(&h, &two, 1);
    for (int i = 0; i < 5; i++) { size_t l = enc_point(S, pts[i], e); h_update(&h, e, l); }
    h_update(&h, ad, adlen); h_update(&h, &zero, 1); h_final(&h, d);
    fe m; f_from_bytes_mod(S->C->Fr, &m, d, (size_t)S->clen, 1); f_to_raw(S->C->Fr, c_raw, &m);
}
static void suite_point_to_hash(const suite_t *S, const aff *P, uint8_t *out) {
    uint8_t three = 0x03, zero = 0x00, e[33]; hasher h; h_init(&h, S->is512);
    h_update(&h, S->suite_id, S->suite_id_len); h_update(&h, &three, 1);
    size_t l = enc_point(S, P, e); h_update(&h, e, l); h_update(&h, &zero, 1); h_final(&h, out);
}
static void pedersen_blinding(const suite_t *S, fe *b_raw, const fe *sk_raw, const aff *I, const uint8_t *ad, size_t adlen) {
    uint8_t cc = 0xCC, zero = 0x00, e[33], d[64]; hasher h; h_init(&h, S->is512);
    h_update(&h, S->suite_id, S->suite_id_len); h_update(&h, &cc, 1);
    enc_scalar(S, sk_raw, e); h_update(&h, e, 32);
    size_t l = enc_point(S, I, e); h_update(&h, e, l); h_update(&h, ad, adlen); h_update(&h, &zero, 1); h_final(&h, d);
    fe m; f_from_bytes_mod(S->C->Fr, &m, d, S->is512 ? 64 : 32, 1); f_to_raw(S->C->Fr, b_raw, &m);
}
/* r = a + b*c mod order, canonical in/out */
static void sc_muladd(const fctx *Fr, fe *r, const fe *a, const fe *b, const fe *c) {
    fe am, bm, cm; f_from_raw(Fr, &am, a); f_from_raw(Fr, &bm, b); f_from_raw(Fr, &cm, c);
    f_mul(Fr, &bm, &bm, &cm); f_add(Fr, &am, &am, &bm); f_to_raw(Fr, r, &am);
}

/* ======================================================================================
 * IETF and Pedersen VRF, one item  (ietf::{Prover,Verifier}, pedersen::{Prover,Verifier}; A.9, A.10)
 * ====================================================================================== */
static void ietf_prove_one(const suite_t *S, const fe *sk, const aff *I, const aff *O, const uint8_t *ad, size_t adlen, fe *c, fe *s) {
    const curve *C = S->C; aff Y, kG, kI; fe k;
    aff_mul(C, &Y, sk, &C->G);
    suite_nonce(S, &k, sk, I);
    aff_mul(C, &kG, &k, &C->G); aff_mul(C, &kI, &k, I);
    const aff *pts[5] = {&Y, I, O, &kG, &kI};
    suite_challenge(S, c, pts, ad, adlen);
    sc_muladd(C->Fr, s, &k, c, sk);
}
static int ietf_verify_one(const suite_t *S, const aff *Y, const aff *I, const aff *O, const fe *c, const fe *s, const uint8_t *ad, size_t adlen) {
    const curve *C = S->C; proj pG, pY, pI, pO, a, b; aff U, V;
    pt_from_aff(C, &pG, &C->G); pt_from_aff(C, &pY, Y); pt_from_aff(C, &pI, I); pt_from_aff(C, &pO, O);
    pt_mul(C, &a, s->v, 4, &pG); pt_mul(C, &b, c->v, 4, &pY); pt_neg(C, &b, &b); pt_add(C, &a, &a, &b); pt_to_aff(C, &U, &a);
    pt_mul(C, &a, s->v, 4, &pI); pt_mul(C, &b, c->v, 4, &pO); pt_neg(C, &b, &b); pt_add(C, &a, &a, &b); pt_to_aff(C, &V, &a);
    if (!C->is_te && (U.inf || V.inf || Y->inf || I->inf || O->inf)) return 0;     /* identity is not encodable in SEC1-compressed form */
    const aff *pts[5] = {Y, I, O, &U, &V}; fe c2;
    suite_challenge(S, &c2, pts, ad, adlen);
    return f_eq(C->Fr, &c2, c);
}
typedef struct { aff Yb, R, Ok; fe s, sb; } ped_proof;
static void pedersen_prove_one(const suite_t *S, const fe *sk, const aff *I, const aff *O, const uint8_t *ad, size_t adlen, ped_proof *pr, fe *blinding) {
    const curve *C = S->C; fe b, k, kb, c; proj pG, pB, pI, t, u;
    pedersen_blinding(S, &b, sk, I, ad, adlen);
    suite_nonce(S, &k, sk, I); suite_nonce(S, &kb, &b, I);
    pt_from_aff(C, &pG, &C->G); pt_from_aff(C, &pB, &S->B); pt_from_aff(C, &pI, I);
    pt_mul(C, &t, sk->v, 4, &pG); pt_mul(C, &u, b.v, 4, &pB); pt_add(C, &t, &t, &u); pt_to_aff(C, &pr->Yb, &t);
    pt_mul(C, &t, k.v, 4, &pG); pt_mul(C, &u, kb.v, 4, &pB); pt_add(C, &t, &t, &u); pt_to_aff(C, &pr->R, &t);
    pt_mul(C, &t, k.v, 4, &pI); pt_to_aff(C, &pr->Ok, &t);
    const aff *pts[5] = {&pr->Yb, I, O, &pr->R, &pr->Ok};
    suite_challenge(S, &c, pts, ad, adlen);
    sc_muladd(C->Fr, &pr->s, &k, &c, sk); sc_muladd(C->Fr, &pr->sb, &kb, &c, &b);
    *blinding = b;
}
static int pedersen_verify_one(const suite_t *S, const aff *I, const aff *O, const ped_proof *pr, const uint8_t *ad, size_t adlen) {
    const curve *C = S->C; fe c; proj pG, pB, pI, pO, pYb, pR, pOk, l, r, t; aff la, ra;
    if (!C->is_te && (I->inf || O->inf || pr->Yb.inf || pr->R.inf || pr->Ok.inf)) return 0;
    const aff *pts[5] = {&pr->Yb, I, O, &pr->R, &pr->Ok};
    suite_challenge(S, &c, pts, ad, adlen);
    pt_from_aff(C, &pG, &C->G); pt_from_aff(C, &pB, &S->B); pt_from_aff(C, &pI, I); pt_from_aff(C, &pO, O);
    pt_from_aff(C, &pYb, &pr->Yb); pt_from_aff(C, &pR, &pr->R); pt_from_aff(C, &pOk, &pr->Ok);
    pt_mul(C, &t, c.v, 4, &pO); pt_add(C, &l, &pOk, &t); pt_mul(C, &r, pr->s.v, 4, &pI);
    pt_to_aff(C, &la, &l); pt_to_aff(C, &ra, &r); if (!aff_eq(C, &la, &ra)) return 0;
    pt_mul(C, &t, c.v, 4, &pYb); pt_add(C, &l, &pR, &t);
    pt_mul(C, &r, pr->s.v, 4, &pG); pt_mul(C, &t, pr->sb.v, 4, &pB); pt_add(C, &r, &r, &t);
    pt_to_aff(C, &la, &l); pt_to_aff(C, &ra, &r); return aff_eq(C, &la, &ra);
}

/* ======================================================================================
 * Batch drivers: contiguous index ranges over pthreads (the reference's rayon-style data parallelism)
 * ====================================================================================== */
typedef void (*item_fn)(void *ctx, size_t i);
typedef struct { item_fn fn; void *ctx; size_t lo, hi; } range_job;
static void *range_worker(void *p) { range_job *j = p; for (size_t i = j->lo; i < j->hi; i++) j->fn(j->ctx, i); return NULL; }
static void parallel_for(size_t n, int nthreads, item_fn fn, void *ctx) {
    pthread_once(&init_once, init_all);
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    if (nthreads == 1) { for (size_t i = 0; i < n; i++) fn(ctx, i); return; }
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads); range_job *jobs = malloc(sizeof(range_job) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (range_job){fn, ctx, n * (size_t)t / (size_t)nthreads, n * (size_t)(t + 1) / (size_t)nthreads};
        pthread_create(&th[t], NULL, range_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

typedef struct {
    const suite_t *S; const uint8_t *a, *b, *c, *d, *e, *ad; const uint64_t *off; uint8_t *o1, *o2, *o3;
    uint8_t *st;   /* optional per-item Result<(), Error>: 0 Ok, 1 Error::VerificationFailure, 2 Error::InvalidData */
} bctx;
#define ST_OK 0
#define ST_VERIFICATION_FAILURE 1
#define ST_INVALID_DATA 2
#define SET_ST(x, i, v) do { if ((x)->st) (x)->st[i] = (uint8_t)(v); } while (0)
#define AD_PTR(x, i) ((x)->ad && (x)->off ? (x)->ad + (x)->off[i] : (const uint8_t *)"")
#define AD_LEN(x, i) ((x)->ad && (x)->off ? (size_t)((x)->off[(i) + 1] - (x)->off[i]) : (size_t)0)

static void it_from_seed(void *p, size_t i) {
    bctx *x = p; const suite_t *S = x->S; uint8_t d[64]; hasher h; h_init(&h, S->is512);
    h_update(&h, x->a + x->off[i], (size_t)(x->off[i + 1] - x->off[i])); h_final(&h, d);
    fe m, sk; f_from_bytes_mod(S->C->Fr, &m, d, S->is512 ? 64 : 32, 0); f_to_raw(S->C->Fr, &sk, &m);
    store_scalar(&sk, x->o1 + 32 * i);
    if (x->o2) { aff Y; aff_mul(S->C, &Y, &sk, &S->C->G); store_point(S->C, &Y, x->o2 + 64 * i); }
}
void oracle_secret_from_seed_batch(int suite, size_t n, const uint8_t *seeds, const uint64_t *seed_off, uint8_t *out_sk, uint8_t *out_pk, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = seeds; x.off = seed_off; x.o1 = out_sk; x.o2 = out_pk; parallel_for(n, nthreads, it_from_seed, &x);
}
static void it_enc(void *p, size_t i) {
    bctx *x = p; aff A; size_t L = pt_len(x->S);
    if (!load_point(x->S->C, &A, x->a + 64 * i) || (!x->S->C->is_te && A.inf)) { memset(x->o1 + L * i, 0, L); return; }
    enc_point(x->S, &A, x->o1 + L * i);
}
void oracle_point_encode_batch(int suite, size_t n, const uint8_t *pts, uint8_t *out_enc, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = pts; x.o1 = out_enc; parallel_for(n, nthreads, it_enc, &x);
}
static void it_dec(void *p, size_t i) {
    bctx *x = p; aff A; size_t L = pt_len(x->S); int ok = dec_point(x->S, &A, x->a + L * i);
    x->o2[i] = (uint8_t)ok; if (ok) store_point(x->S->C, &A, x->o1 + 64 * i); else memset(x->o1 + 64 * i, 0, 64);
}
void oracle_point_decode_batch(int suite, size_t n, const uint8_t *enc, uint8_t *out_pts, uint8_t *out_ok, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = enc; x.o1 = out_pts; x.o2 = out_ok; parallel_for(n, nthreads, it_dec, &x);
}
static void it_h2c(void *p, size_t i) {
    bctx *x = p; aff A; int ok = data_to_point(x->S, &A, x->a + x->off[i], (size_t)(x->off[i + 1] - x->off[i]));
    x->o2[i] = (uint8_t)ok; if (ok) store_point(x->S->C, &A, x->o1 + 64 * i); else memset(x->o1 + 64 * i, 0, 64);
}
void oracle_data_to_point_batch(int suite, size_t n, const uint8_t *data, const uint64_t *data_off, uint8_t *out_pts, uint8_t *out_ok, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = data; x.off = data_off; x.o1 = out_pts; x.o2 = out_ok; parallel_for(n, nthreads, it_h2c, &x);
}
static void it_output(void *p, size_t i) {
    bctx *x = p; const curve *C = x->S->C; fe sk; aff I, O; load_scalar(C, &sk, x->a + 32 * i);
    if (!load_point(C, &I, x->b + 64 * i)) { memset(x->o1 + 64 * i, 0, 64); return; }
    aff_mul(C, &O, &sk, &I); store_point(C, &O, x->o1 + 64 * i);
}
void oracle_output_batch(int suite, size_t n, const uint8_t *sk, const uint8_t *input, uint8_t *out_output, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = sk; x.b = input; x.o1 = out_output; parallel_for(n, nthreads, it_output, &x);
}
static void it_pt_hash(void *p, size_t i) {
    bctx *x = p; aff A; size_t L = x->S->is512 ? 64 : 32;
    if (!load_point(x->S->C, &A, x->a + 64 * i) || (!x->S->C->is_te && A.inf)) { memset(x->o1 + L * i, 0, L); return; }
    suite_point_to_hash(x->S, &A, x->o1 + L * i);
}
void oracle_point_to_hash_batch(int suite, size_t n, const uint8_t *pts, uint8_t *out_hash, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = pts; x.o1 = out_hash; parallel_for(n, nthreads, it_pt_hash, &x);
}
static void it_nonce(void *p, size_t i) {
    bctx *x = p; const curve *C = x->S->C; fe sk, k; aff I; load_scalar(C, &sk, x->a + 32 * i); load_point(C, &I, x->b + 64 * i);
    suite_nonce(x->S, &k, &sk, &I); store_scalar(&k, x->o1 + 32 * i);
}
void oracle_nonce_batch(int suite, size_t n, const uint8_t *sk, const uint8_t *input, uint8_t *out_k, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = sk; x.b = input; x.o1 = out_k; parallel_for(n, nthreads, it_nonce, &x);
}
static void it_ietf_prove(void *p, size_t i) {
    bctx *x = p; const curve *C = x->S->C; fe sk, c, s; aff I, O;
    load_scalar(C, &sk, x->a + 32 * i); load_point(C, &I, x->b + 64 * i); load_point(C, &O, x->c + 64 * i);
    ietf_prove_one(x->S, &sk, &I, &O, AD_PTR(x, i), AD_LEN(x, i), &c, &s);
    store_scalar(&c, x->o1 + 32 * i); store_scalar(&s, x->o2 + 32 * i);
}
void oracle_ietf_prove_batch(int suite, size_t n, const uint8_t *sk, const uint8_t *input, const uint8_t *output, const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_c, uint8_t *out_s, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = sk; x.b = input; x.c = output; x.ad = ad; x.off = ad_off; x.o1 = out_c; x.o2 = out_s;
    parallel_for(n, nthreads, it_ietf_prove, &x);
}
static void it_ietf_verify(void *p, size_t i) {
    bctx *x = p; const curve *C = x->S->C; fe c, s; aff Y, I, O;
    int ok = load_point(C, &Y, x->a + 64 * i) & load_point(C, &I, x->b + 64 * i) & load_point(C, &O, x->c + 64 * i);
    ok = ok && aff_on_curve(C, &Y) && aff_on_curve(C, &I) && aff_on_curve(C, &O);   /* typed Public/Input/Output are on-curve by construction */
    load_scalar(C, &c, x->d + 32 * i); load_scalar(C, &s, x->e + 32 * i);
    /* values that no typed Public / Input / Output can hold (non-canonical, off the curve, the un-encodable short-Weierstrass
     * identity) are what deserialisation rejects with Error::InvalidData; a proof that does not check is VerificationFailure */
    if (ok && !C->is_te && (Y.inf || I.inf || O.inf)) ok = 0;
    if (!ok) { x->o1[i] = 0; SET_ST(x, i, ST_INVALID_DATA); return; }
    const int good = ietf_verify_one(x->S, &Y, &I, &O, &c, &s, AD_PTR(x, i), AD_LEN(x, i));
    x->o1[i] = (uint8_t)good; SET_ST(x, i, good ? ST_OK : ST_VERIFICATION_FAILURE);
}
void oracle_ietf_verify_status_batch(int suite, size_t n, const uint8_t *pk, const uint8_t *input, const uint8_t *output, const uint8_t *c, const uint8_t *s, const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, uint8_t *out_status, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = pk; x.b = input; x.c = output; x.d = c; x.e = s; x.ad = ad; x.off = ad_off; x.o1 = out_ok; x.st = out_status;
    parallel_for(n, nthreads, it_ietf_verify, &x);
}
void oracle_ietf_verify_batch(int suite, size_t n, const uint8_t *pk, const uint8_t *input, const uint8_t *output, const uint8_t *c, const uint8_t *s, const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, int nthreads) {
    oracle_ietf_verify_status_batch(suite, n, pk, input, output, c, s, ad, ad_off, out_ok, NULL, nthreads);
}
static void it_ped_prove(void *p, size_t i) {
    bctx *x = p; const curve *C = x->S->C; fe sk, bl; aff I, O; ped_proof pr;
    load_scalar(C, &sk, x->a + 32 * i); load_point(C, &I, x->b + 64 * i); load_point(C, &O, x->c + 64 * i);
    pedersen_prove_one(x->S, &sk, &I, &O, AD_PTR(x, i), AD_LEN(x, i), &pr, &bl);
    uint8_t *o = x->o1 + 256 * i; store_point(C, &pr.Yb, o); store_point(C, &pr.R, o + 64); store_point(C, &pr.Ok, o + 128);
    store_scalar(&pr.s, o + 192); store_scalar(&pr.sb, o + 224); store_scalar(&bl, x->o2 + 32 * i);
}
void oracle_pedersen_prove_batch(int suite, size_t n, const uint8_t *sk, const uint8_t *input, const uint8_t *output, const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_proof, uint8_t *out_blinding, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = sk; x.b = input; x.c = output; x.ad = ad; x.off = ad_off; x.o1 = out_proof; x.o2 = out_blinding;
    parallel_for(n, nthreads, it_ped_prove, &x);
}
static void it_ped_verify(void *p, size_t i) {
    bctx *x = p; const curve *C = x->S->C; aff I, O; ped_proof pr; const uint8_t *pb = x->c + 256 * i;
    int ok = load_point(C, &I, x->a + 64 * i) & load_point(C, &O, x->b + 64 * i) & load_point(C, &pr.Yb, pb) & load_point(C, &pr.R, pb + 64) & load_point(C, &pr.Ok, pb + 128);
    ok = ok && aff_on_curve(C, &I) && aff_on_curve(C, &O) && aff_on_curve(C, &pr.Yb) && aff_on_curve(C, &pr.R) && aff_on_curve(C, &pr.Ok);
    load_scalar(C, &pr.s, pb + 192); load_scalar(C, &pr.sb, pb + 224);
    if (ok && !C->is_te && (I.inf || O.inf || pr.Yb.inf || pr.R.inf || pr.Ok.inf)) ok = 0;
    if (!ok) { x->o1[i] = 0; SET_ST(x, i, ST_INVALID_DATA); return; }
    const int good = pedersen_verify_one(x->S, &I, &O, &pr, AD_PTR(x, i), AD_LEN(x, i));
    x->o1[i] = (uint8_t)good; SET_ST(x, i, good ? ST_OK : ST_VERIFICATION_FAILURE);
}
void oracle_pedersen_verify_status_batch(int suite, size_t n, const uint8_t *input, const uint8_t *output, const uint8_t *proof, const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, uint8_t *out_status, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = input; x.b = output; x.c = proof; x.ad = ad; x.off = ad_off; x.o1 = out_ok; x.st = out_status;
    parallel_for(n, nthreads, it_ped_verify, &x);
}
void oracle_pedersen_verify_batch(int suite, size_t n, const uint8_t *input, const uint8_t *output, const uint8_t *proof, const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, int nthreads) {
    oracle_pedersen_verify_status_batch(suite, n, input, output, proof, ad, ad_off, out_ok, NULL, nthreads);
}

/* ======================================================================================
 * Wire formats: what ark-serialize's CanonicalDeserialize (Compress::Yes, Validate::Yes) accepts for
 * `Public` / `Output` (codec point + on-curve + ark-ec's default `is_in_correct_subgroup_assuming_on_curve`, i.e.
 * mul_bigint(r).is_zero()) and for `ietf::Proof` (c: CHALLENGE_LEN bytes through the codec's scalar_decode, i.e. reduced
 * mod r; s: a canonical scalar, rejected when >= r).  Byte layout per A.9: c || s in codec byte order.  [RECALL] for the
 * two scalar rules; the point rules are arkworks' documented behaviour.
 * signature = point_encode(Output) || c || s   (Bandersnatch 96 B; secp256r1 81 B = RFC 9381 pi_string)
 * ====================================================================================== */
static int in_prime_subgroup(const curve *C, const aff *A) {
    if (C->cof_log2 == 0) return 1;                /* cofactor 1 */
    const fctx *F = C->F; proj p, r; pt_from_aff(C, &p, A);
    pt_mul(C, &r, C->Fr->p.v, C->Fr->n, &p);
    if (!C->is_te) return pt_is_identity(C, &r);   /* short Weierstrass with a cofactor (bandersnatch_sw) */
    /* ark-ec twisted_edwards::Projective::is_zero: x == 0 && y == z && y != 0 && t == 0 */
    return f_is_zero(F, &r.X) && f_eq(F, &r.Y, &r.Z) && !f_is_zero(F, &r.Y) && f_is_zero(F, &r.T);
}
static int dec_point_checked(const suite_t *S, aff *A, const uint8_t *in) {
    return dec_point(S, A, in) && aff_on_curve(S->C, A) && in_prime_subgroup(S->C, A);
}
static void it_subgroup(void *p, size_t i) {
    bctx *x = p; aff A; int ok = load_point(x->S->C, &A, x->a + 64 * i) && aff_on_curve(x->S->C, &A) && in_prime_subgroup(x->S->C, &A);
    x->o1[i] = (uint8_t)ok;
}
void oracle_subgroup_check_batch(int suite, size_t n, const uint8_t *pts, uint8_t *out_ok, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = pts; x.o1 = out_ok; parallel_for(n, nthreads, it_subgroup, &x);
}
static void it_dec_checked(void *p, size_t i) {
    bctx *x = p; aff A; size_t L = pt_len(x->S); int ok = dec_point_checked(x->S, &A, x->a + L * i);
    x->o2[i] = (uint8_t)ok; if (ok) store_point(x->S->C, &A, x->o1 + 64 * i); else memset(x->o1 + 64 * i, 0, 64);
}
void oracle_point_decode_checked_batch(int suite, size_t n, const uint8_t *enc, uint8_t *out_pts, uint8_t *out_ok, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = enc; x.o1 = out_pts; x.o2 = out_ok; parallel_for(n, nthreads, it_dec_checked, &x);
}
int oracle_ietf_signature_len(int suite) { const suite_t *S = get_suite(suite); return pt_len(S) + S->clen + 32; }
/* c (canonical raw) -> CHALLENGE_LEN bytes in codec order */
static void enc_challenge(const suite_t *S, const fe *c_raw, uint8_t *out) {
    uint8_t t[32]; fe k = *c_raw; store_scalar(&k, t);
    if (S->sec1) for (int i = 0; i < S->clen; i++) out[i] = t[S->clen - 1 - i]; else memcpy(out, t, (size_t)S->clen);
}
static void it_sign_wire(void *p, size_t i) {
    bctx *x = p; const suite_t *S = x->S; const curve *C = S->C; size_t PL = pt_len(S), SL = PL + (size_t)S->clen + 32;
    uint8_t *sig = x->o1 + SL * i; fe sk, c, s; aff I, O;
    load_scalar(C, &sk, x->a + 32 * i);
    if (!data_to_point(S, &I, x->b + x->off[i], (size_t)(x->off[i + 1] - x->off[i]))) { memset(sig, 0, SL); x->o2[i] = 0; return; }
    aff_mul(C, &O, &sk, &I);
    const uint8_t *ad = x->c ? x->c + ((const uint64_t *)x->d)[i] : (const uint8_t *)"";
    size_t adlen = x->c ? (size_t)(((const uint64_t *)x->d)[i + 1] - ((const uint64_t *)x->d)[i]) : 0;
    ietf_prove_one(S, &sk, &I, &O, ad, adlen, &c, &s);
    enc_point(S, &O, sig); enc_challenge(S, &c, sig + PL); enc_scalar(S, &s, sig + PL + S->clen);
    x->o2[i] = 1;
}
/* Secret::from(sk): Input::new(data) -> Secret::output -> ietf::Prover::prove -> serialise */
void oracle_ietf_sign_wire_batch(int suite, size_t n, const uint8_t *sk, const uint8_t *data, const uint64_t *data_off,
                                 const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_sig, uint8_t *out_ok, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = sk; x.b = data; x.off = data_off; x.c = ad; x.d = (const uint8_t *)ad_off; x.o1 = out_sig; x.o2 = out_ok;
    parallel_for(n, nthreads, it_sign_wire, &x);
}
static void it_verify_wire(void *p, size_t i) {
    bctx *x = p; const suite_t *S = x->S; const curve *C = S->C; size_t PL = pt_len(S), SL = PL + (size_t)S->clen + 32;
    const uint8_t *sig = x->e + SL * i; aff Y, I, O; fe c, s, m;
    x->o1[i] = 0; if (x->o2) memset(x->o2 + (size_t)(S->is512 ? 64 : 32) * i, 0, S->is512 ? 64 : 32);
    SET_ST(x, i, ST_INVALID_DATA);                                                       /* every early return below is a failed deserialisation */
    if (!dec_point_checked(S, &Y, x->a + PL * i)) return;                                /* Public::deserialize_compressed */
    if (!dec_point_checked(S, &O, sig)) return;                                          /* Output */
    if (!C->is_te && (Y.inf || O.inf)) return;
    uint8_t le[32] = {0};                                                                /* Proof.c: clen bytes, reduced mod r */
    for (int j = 0; j < S->clen; j++) le[j] = S->sec1 ? sig[PL + S->clen - 1 - j] : sig[PL + j];
    f_from_bytes_mod(C->Fr, &m, le, 32, 0); f_to_raw(C->Fr, &c, &m);
    for (int j = 0; j < 32; j++) le[j] = S->sec1 ? sig[PL + S->clen + 31 - j] : sig[PL + S->clen + j];   /* Proof.s: canonical */
    if (!f_from_le_canonical(C->Fr, &m, le)) return;
    f_to_raw(C->Fr, &s, &m);
    if (!data_to_point(S, &I, x->b + x->off[i], (size_t)(x->off[i + 1] - x->off[i]))) return;   /* Input::new */
    const uint8_t *ad = x->c ? x->c + ((const uint64_t *)x->d)[i] : (const uint8_t *)"";
    size_t adlen = x->c ? (size_t)(((const uint64_t *)x->d)[i + 1] - ((const uint64_t *)x->d)[i]) : 0;
    if (!ietf_verify_one(S, &Y, &I, &O, &c, &s, ad, adlen)) { SET_ST(x, i, ST_VERIFICATION_FAILURE); return; }
    SET_ST(x, i, ST_OK);
    x->o1[i] = 1; if (x->o2) suite_point_to_hash(S, &O, x->o2 + (size_t)(S->is512 ? 64 : 32) * i);
}
/* Public::deserialize + Input::new(data) + Output/Proof::deserialize + ietf::Verifier::verify (+ Output::hash for accepted items) */
void oracle_ietf_verify_wire_status_batch(int suite, size_t n, const uint8_t *pk_enc, const uint8_t *data, const uint64_t *data_off, const uint8_t *sig,
                                          const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, uint8_t *out_hash, uint8_t *out_status, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = pk_enc; x.b = data; x.off = data_off; x.e = sig; x.c = ad; x.d = (const uint8_t *)ad_off; x.o1 = out_ok; x.o2 = out_hash; x.st = out_status;
    parallel_for(n, nthreads, it_verify_wire, &x);
}
void oracle_ietf_verify_wire_batch(int suite, size_t n, const uint8_t *pk_enc, const uint8_t *data, const uint64_t *data_off, const uint8_t *sig,
                                   const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, uint8_t *out_hash, int nthreads) {
    oracle_ietf_verify_wire_status_batch(suite, n, pk_enc, data, data_off, sig, ad, ad_off, out_ok, out_hash, NULL, nthreads);
}

/* pedersen wire form: point_encode(Output) || point_encode(pk_com) || point_encode(r) || point_encode(ok) || s || sb
 * (`Output` followed by `pedersen::Proof`'s CanonicalSerialize, A.10); the blinding factor stays with the prover. */
int oracle_pedersen_signature_len(int suite) { const suite_t *S = get_suite(suite); return 4 * pt_len(S) + 64; }
static void it_ped_sign_wire(void *p, size_t i) {
    bctx *x = p; const suite_t *S = x->S; const curve *C = S->C; size_t PL = pt_len(S), SL = 4 * PL + 64;
    uint8_t *sig = x->o1 + SL * i; fe sk, bl; aff I, O; ped_proof pr;
    load_scalar(C, &sk, x->a + 32 * i);
    if (!data_to_point(S, &I, x->b + x->off[i], (size_t)(x->off[i + 1] - x->off[i]))) { memset(sig, 0, SL); memset(x->o3 + 32 * i, 0, 32); x->o2[i] = 0; return; }
    aff_mul(C, &O, &sk, &I);
    const uint8_t *ad = x->c ? x->c + ((const uint64_t *)x->d)[i] : (const uint8_t *)"";
    size_t adlen = x->c ? (size_t)(((const uint64_t *)x->d)[i + 1] - ((const uint64_t *)x->d)[i]) : 0;
    pedersen_prove_one(S, &sk, &I, &O, ad, adlen, &pr, &bl);
    enc_point(S, &O, sig); enc_point(S, &pr.Yb, sig + PL); enc_point(S, &pr.R, sig + 2 * PL); enc_point(S, &pr.Ok, sig + 3 * PL);
    enc_scalar(S, &pr.s, sig + 4 * PL); enc_scalar(S, &pr.sb, sig + 4 * PL + 32);
    store_scalar(&bl, x->o3 + 32 * i);
    x->o2[i] = 1;
}
void oracle_pedersen_sign_wire_batch(int suite, size_t n, const uint8_t *sk, const uint8_t *data, const uint64_t *data_off,
                                     const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_sig, uint8_t *out_blinding, uint8_t *out_ok, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.a = sk; x.b = data; x.off = data_off; x.c = ad; x.d = (const uint8_t *)ad_off; x.o1 = out_sig; x.o2 = out_ok; x.o3 = out_blinding;
    parallel_for(n, nthreads, it_ped_sign_wire, &x);
}
static int dec_scalar_canonical(const suite_t *S, fe *k_raw, const uint8_t *in) {
    uint8_t le[32]; fe m;
    for (int j = 0; j < 32; j++) le[j] = S->sec1 ? in[31 - j] : in[j];
    if (!f_from_le_canonical(S->C->Fr, &m, le)) return 0;
    f_to_raw(S->C->Fr, k_raw, &m); return 1;
}
static void it_ped_verify_wire(void *p, size_t i) {
    bctx *x = p; const suite_t *S = x->S; size_t PL = pt_len(S), SL = 4 * PL + 64;
    const uint8_t *sig = x->e + SL * i; aff I, O; ped_proof pr;
    x->o1[i] = 0; SET_ST(x, i, ST_INVALID_DATA);
    if (!dec_point_checked(S, &O, sig) || !dec_point_checked(S, &pr.Yb, sig + PL) || !dec_point_checked(S, &pr.R, sig + 2 * PL) || !dec_point_checked(S, &pr.Ok, sig + 3 * PL)) return;
    if (!dec_scalar_canonical(S, &pr.s, sig + 4 * PL) || !dec_scalar_canonical(S, &pr.sb, sig + 4 * PL + 32)) return;
    if (!data_to_point(S, &I, x->b + x->off[i], (size_t)(x->off[i + 1] - x->off[i]))) return;
    const uint8_t *ad = x->c ? x->c + ((const uint64_t *)x->d)[i] : (const uint8_t *)"";
    size_t adlen = x->c ? (size_t)(((const uint64_t *)x->d)[i + 1] - ((const uint64_t *)x->d)[i]) : 0;
    if (!S->C->is_te && (I.inf || O.inf || pr.Yb.inf || pr.R.inf || pr.Ok.inf)) return;
    x->o1[i] = (uint8_t)pedersen_verify_one(S, &I, &O, &pr, ad, adlen);
    SET_ST(x, i, x->o1[i] ? ST_OK : ST_VERIFICATION_FAILURE);
}
void oracle_pedersen_verify_wire_status_batch(int suite, size_t n, const uint8_t *data, const uint64_t *data_off, const uint8_t *sig,
                                              const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, uint8_t *out_status, int nthreads) {
    bctx x = {0}; x.S = get_suite(suite); x.b = data; x.off = data_off; x.e = sig; x.c = ad; x.d = (const uint8_t *)ad_off; x.o1 = out_ok; x.st = out_status;
    parallel_for(n, nthreads, it_ped_verify_wire, &x);
}
void oracle_pedersen_verify_wire_batch(int suite, size_t n, const uint8_t *data, const uint64_t *data_off, const uint8_t *sig,
                                       const uint8_t *ad, const uint64_t *ad_off, uint8_t *out_ok, int nthreads) {
    oracle_pedersen_verify_wire_status_batch(suite, n, data, data_off, sig, ad, ad_off, out_ok, NULL, nthreads);
}

/* ======================================================================================
 * BLS12-381 G1 MSM  (ark-ec VariableBaseMSM::msm -> msm_bigint: Pippenger with
 * c = 3 for n < 32, else ln(n) + 2; per-window bucket accumulation + running sum; windows
 * combined MSB-first with c doublings).  rayon parallelises over windows; so do we.
 * ====================================================================================== */
static int ln_without_floats(size_t a) { int lg = 0; while ((a >> lg) > 1) lg++; return lg * 69 / 100; }
typedef struct { size_t n; const aff *bases; const fe *sc; int c, nwin; proj *win; } msm_ctx;
static void it_msm_window(void *p, size_t w) {
    msm_ctx *m = p; const curve *C = &C_G1; size_t nb = ((size_t)1 << m->c) - 1;
    proj *buckets = malloc(sizeof(proj) * nb); for (size_t i = 0; i < nb; i++) pt_identity(C, &buckets[i]);
    size_t start = w * (size_t)m->c; proj res; pt_identity(C, &res);
    for (size_t i = 0; i < m->n; i++) {
        const fe *s = &m->sc[i]; size_t limb = start / 64, sh = start % 64;
        u64 v = s->v[limb] >> sh; if (sh && limb + 1 < 4) v |= s->v[limb + 1] << (64 - sh);
        v &= nb;
        if (v) { proj q; pt_from_aff(C, &q, &m->bases[i]); pt_add(C, &buckets[v - 1], &buckets[v - 1], &q); }
    }
    proj run; pt_identity(C, &run);
    for (size_t i = nb; i-- > 0;) { pt_add(C, &run, &run, &buckets[i]); pt_add(C, &res, &res, &run); }
    free(buckets); m->win[w] = res;
}
void oracle_msm_g1(size_t n, const uint8_t *bases, const uint8_t *scalars, int n_columns, uint8_t *out, int nthreads) {
    pthread_once(&init_once, init_all);
    const curve *C = &C_G1;
    aff *B = malloc(sizeof(aff) * (n ? n : 1)); fe *S = malloc(sizeof(fe) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) load_point(C, &B[i], bases + 96 * i);
    int c = n < 32 ? 3 : ln_without_floats(n) + 2; int nwin = (255 + c - 1) / c;
    proj *win = malloc(sizeof(proj) * (size_t)nwin);
    for (int col = 0; col < n_columns; col++) {
        for (size_t i = 0; i < n; i++) { fe m; f_from_bytes_mod(&F_BLSFR, &m, scalars + 32 * ((size_t)col * n + i), 32, 0); f_to_raw(&F_BLSFR, &S[i], &m); }
        msm_ctx m = {n, B, S, c, nwin, win};
        parallel_for((size_t)nwin, nthreads, it_msm_window, &m);
        proj total = win[nwin - 1];
        for (int w = nwin - 2; w >= 0; w--) { for (int k = 0; k < c; k++) pt_double(C, &total, &total); pt_add(C, &total, &total, &win[w]); }
        aff R; pt_to_aff(C, &R, &total); store_point(C, &R, out + 96 * (size_t)col);
    }
    free(B); free(S); free(win);
}
typedef struct { const uint8_t *sc; uint8_t *out; } g1mul_ctx;
static void it_g1mul(void *p, size_t i) {
    g1mul_ctx *x = p; fe m, k; f_from_bytes_mod(&F_BLSFR, &m, x->sc + 32 * i, 32, 0); f_to_raw(&F_BLSFR, &k, &m);
    aff R; aff_mul(&C_G1, &R, &k, &C_G1.G); store_point(&C_G1, &R, x->out + 96 * i);
}
void oracle_g1_mul_gen_batch(size_t n, const uint8_t *scalars, uint8_t *out_pts, int nthreads) {
    g1mul_ctx x = {scalars, out_pts}; parallel_for(n, nthreads, it_g1mul, &x);
}
