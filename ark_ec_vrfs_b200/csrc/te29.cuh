// Bandersnatch point arithmetic on the unsaturated 9x29-bit field (f29.cuh): the same hwcd formulas as
// te.cuh, with lazy signed limbs.  Bound bookkeeping (units: p for values, powers of two for |limbs|):
//   every stored coordinate is an f29_mul / f29_sqr output: normalised, |value| < 2p.
//   add:  A,B,C,D = products of stored coords            (N x N)             < 1.06
//         s1 = norm(X1+Y1) < 4 ; s2 = +-X2 + Y2 (|limb| < 2^30)   s1*s2      (N x 2^30)
//         E = norm(s1*s2 - A - B) < 3.4 ; F = D - C (2^29) < 2.2 ; G = D + C (2^30) < 2.2 ; H = norm(B + 5A) < 6.4
//         X3 = E*F, Y3 = H*G, T3 = E*H, Z3 = F*G :  |a||b| <= 22 p^2 < 70 p^2, limb products <= 2^59   -> outputs < 1.4
//   dbl:  A,B,Cz = squares of stored coords < 1.06 ; s = norm(X+Y) ; E = norm(s^2 - A - B) < 3.3
//         S5 = norm(5A) < 5.3 ; G = B - S5 (2^29) < 6.4 ; H = -(S5 + B) (2^30) < 6.4 ; F = norm(G - 2Cz) < 8.5
//         X3 = E*F (28), Y3 = G*H (41), Z3 = F*G (54), T3 = E*H (21) p^2  -> outputs < 1.8
#pragma once
#include "f29.cuh"
#include "te.cuh"

namespace vrfs {

struct alignas(16) TE29Point { F29 X, Y, Z, T; };
struct alignas(16) TE29Cached { F29 X, Y, Z, dT; };                 // 144 bytes
struct alignas(16) TE29AffCached { F29 x, y, dt; uint32_t pad; };   // 112 bytes

HD_INLINE void te29_set_identity(TE29Point& P) { P.X = f29_zero(); P.Y = f29_one(); P.Z = f29_one(); P.T = f29_zero(); }
HD_INLINE void te29_to_cached(TE29Cached& r, const TE29Point& P) { r.X = P.X; r.Y = P.Y; r.Z = P.Z; r.dT = f29_mul(P.T, f29c<F29Consts::D>()); }

HD_NOINLINE void te29_add_cached(TE29Point* r, const TE29Point* p, const TE29Cached* q, bool negate) {
  F29 qX = f29_cneg(q->X, negate), qdT = f29_cneg(q->dT, negate);
  F29 A = f29_mul(p->X, qX), B = f29_mul(p->Y, q->Y), C = f29_mul(p->T, qdT), D = f29_mul(p->Z, q->Z);
  F29 s1 = f29_norm(f29_add(p->X, p->Y)), s2 = f29_add(qX, q->Y);
  F29 E = f29_norm(f29_sub(f29_sub(f29_mul(s1, s2), A), B));
  F29 F = f29_sub(D, C), G = f29_add(D, C), H = f29_norm_5a_plus_b(A, B);
  r->X = f29_mul(E, F); r->Y = f29_mul(H, G); r->T = f29_mul(E, H); r->Z = f29_mul(F, G);
}
HD_NOINLINE void te29_madd(TE29Point* r, const TE29Point* p, const TE29AffCached* q, bool negate) {
  F29 qx = f29_cneg(q->x, negate), qdt = f29_cneg(q->dt, negate);
  F29 A = f29_mul(p->X, qx), B = f29_mul(p->Y, q->y), C = f29_mul(p->T, qdt), D = p->Z;
  F29 s1 = f29_norm(f29_add(p->X, p->Y)), s2 = f29_add(qx, q->y);
  F29 E = f29_norm(f29_sub(f29_sub(f29_mul(s1, s2), A), B));
  F29 F = f29_sub(D, C), G = f29_add(D, C), H = f29_norm_5a_plus_b(A, B);
  r->X = f29_mul(E, F); r->Y = f29_mul(H, G); r->T = f29_mul(E, H); r->Z = f29_mul(F, G);
}
HD_NOINLINE void te29_dbl(TE29Point* r, const TE29Point* p, bool want_t) {
  F29 A = f29_sqr(p->X), B = f29_sqr(p->Y), Cz = f29_sqr(p->Z);
  F29 s = f29_norm(f29_add(p->X, p->Y));
  F29 E = f29_norm(f29_sub(f29_sub(f29_sqr(s), A), B));
  F29 S5 = f29_norm_5a_plus_b(A, f29_zero());
  F29 G = f29_sub(B, S5), H = f29_neg(f29_add(S5, B)), F = f29_norm(f29_sub(G, f29_dbl(Cz)));
  r->X = f29_mul(E, F); r->Y = f29_mul(G, H); r->Z = f29_mul(F, G);
  if (want_t) r->T = f29_mul(E, H);
}
// GLV endomorphism (see band_endo in te.cuh)
HD_NOINLINE void te29_endo(TE29Point* r, const TE29Point* p) {
  F29 b = f29c<F29Consts::ENDO_B>(), c = f29c<F29Consts::ENDO_C>();
  F29 yy = f29_sqr(p->Y), zz = f29_sqr(p->Z), xy = f29_mul(p->X, p->Y), bzz = f29_mul(b, zz);
  F29 f = f29_mul(c, f29_sub(zz, yy)), g = f29_mul(b, f29_add(yy, bzz)), h = f29_sub(yy, bzz);
  bool degenerate = f29_is_zero(p->X);
  TE29Point id; te29_set_identity(id);
  r->X = f29_select(degenerate, id.X, f29_mul(f, h));
  r->Y = f29_select(degenerate, id.Y, f29_mul(g, xy));
  r->Z = f29_select(degenerate, id.Z, f29_mul(xy, h));
  r->T = f29_select(degenerate, id.T, f29_mul(f, g));
}

}  // namespace vrfs
