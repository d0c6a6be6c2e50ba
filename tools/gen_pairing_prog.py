#!/usr/bin/env python3
"""Compile the BLS12-381 pairing product check into LANE-PARALLEL straight-line programs for csrc/pairing_coop.cuh and emit
ark_ec_vrfs_b200/csrc/gen/pairing_prog.cuh.

Why: one product of two pairings is ~19 000 dependent-looking 12-limb field products; on one thread that is 27 ms of a B200
(the KZG check of SURVEY 8(f)3 ends in exactly one such product).  The arithmetic has plenty of parallelism at the F_q level
(an F_q12 product is 54 independent F_q products between two thin layers of additions), so this script

  1. traces the formulas of csrc/pairing.cuh (same tower, same projective Miller steps, same x-chain for the hard part) over
     symbolic F_q values into a DAG of {MUL, ADD, SUB, HALF, INV} nodes with common subexpressions merged,
  2. list-schedules every SEGMENT of the computation (one Miller iteration, one cyclotomic squaring, ...) into steps of at most
     LANES independent operations (critical-path priority; multiplications first, cheap linear operations fill the idle lanes),
  3. allocates the values to a register file of F_q elements in shared memory (a register is only reused in a later step than
     its last read, so one barrier per step is the only synchronisation the interpreter needs), and
  4. emits the step tables plus the MACRO program (the order in which the kernel runs the segments: 63 Miller iterations, the
     easy part, five exponentiations by the curve parameter, ...).

Before anything is written the emitted tables are EXECUTED here by a Python model of the interpreter and compared with the twin of
the one-thread implementation (tools/gen_pairing_consts.py) and with the naive oracle (oracle/pairing_ref.py).
Build infrastructure only; the library never runs this."""
import os
import random
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import pairing_ref as P  # noqa: E402
import gen_pairing_consts as T  # noqa: E402  (the twin of csrc/pairing.cuh; importing it runs nothing)

OUT = os.path.join(ROOT, "ark_ec_vrfs_b200", "csrc", "gen", "pairing_prog.cuh")
Q = P.Q
LANES = int(os.environ.get("PAIRING_LANES", "64"))   # 64: two warps per product (2.19 ms on a B200; 32 lanes: 2.48 ms)
NPAIRS = 2
SCHED_LIN_FIRST = int(os.environ.get("SCHED_LIN_FIRST", "1"))
MILLER_GROUP = int(os.environ.get("PAIRING_MILLER_GROUP", "4"))     # Miller iterations per segment
CYC_RUNS = [int(x) for x in os.environ.get("PAIRING_CYC_RUNS", "1").split(",")]     # cyclotomic squarings per segment (1 must be present; 1,2,4 measured: no fewer steps)

NOP, MUL, ADD, SUB, HALF, INV = 0, 1, 2, 3, 4, 5
IN = 9                       # pseudo-kind: a value that already sits in a pinned register
COST = {MUL: 10.0, ADD: 1.0, SUB: 1.0, HALF: 1.0, INV: 400.0, IN: 0.0}


# ---- register map (F_q registers of 48 bytes) ---------------------------------------------------------------------------
class Regs:
    def __init__(self):
        self.n = 0
        self.names = {}
    def take(self, name, count=1):
        base = self.n
        self.names[name] = (base, count)
        self.n += count
        return list(range(base, base + count))

REG = Regs()
R_ZERO, = REG.take("ZERO")
R_ONE, = REG.take("ONE")
CONST_F2 = {}                # name -> (reg c0, reg c1), value
for nm, val in (("G6_1_1", T.G6_1[1]), ("G6_2_1", T.G6_2[1]), ("G12_1", T.G12[1]), ("G6_1_2", T.G6_1[2]), ("G6_2_2", T.G6_2[2]), ("G12_2", T.G12[2])):
    CONST_F2[nm] = (REG.take(nm, 2), val)
R_P = [REG.take("P%d" % k, 2) for k in range(NPAIRS)]            # G1 affine x, y
R_Q = [REG.take("Q%d" % k, 4) for k in range(NPAIRS)]            # G2 affine x.c0, x.c1, y.c0, y.c1
R_R = [REG.take("R%d" % k, 6) for k in range(NPAIRS)]            # running G2 point, projective
BLOCKS = {nm: REG.take(nm, 12) for nm in ("F", "X", "B", "M", "A", "BV", "CV", "OUT")}
N_PINNED = REG.n


# ---- tracer -------------------------------------------------------------------------------------------------------------
class Trace:
    def __init__(self):
        self.kind, self.a, self.b, self.reg = [], [], [], []
        self.cse = {}
        self.pinned = {}
    def new(self, kind, a=-1, b=-1, reg=-1):
        if kind in (MUL, ADD) and a > b:
            a, b = b, a
        key = (kind, a, b, reg)
        if key in self.cse:
            return self.cse[key]
        i = len(self.kind)
        self.kind.append(kind); self.a.append(a); self.b.append(b); self.reg.append(reg)
        self.cse[key] = i
        return i
    def pin(self, reg):
        if reg not in self.pinned:
            self.pinned[reg] = self.new(IN, reg=reg)
        return self.pinned[reg]

class E:
    """a symbolic F_q value"""
    __slots__ = ("t", "i")
    def __init__(self, t, i): self.t, self.i = t, i
    def _is(self, reg): return self.t.kind[self.i] == IN and self.t.reg[self.i] == reg
    def __add__(self, o):
        if self._is(R_ZERO): return o
        if o._is(R_ZERO): return self
        return E(self.t, self.t.new(ADD, self.i, o.i))
    def __sub__(self, o):
        if o._is(R_ZERO): return self
        if self.i == o.i: return E(self.t, self.t.pin(R_ZERO))
        return E(self.t, self.t.new(SUB, self.i, o.i))
    def __neg__(self):
        if self._is(R_ZERO): return self
        return E(self.t, self.t.new(SUB, self.t.pin(R_ZERO), self.i))
    def __mul__(self, o):
        if self._is(R_ZERO) or o._is(R_ZERO): return E(self.t, self.t.pin(R_ZERO))
        if self._is(R_ONE): return o
        if o._is(R_ONE): return self
        return E(self.t, self.t.new(MUL, self.i, o.i))
    def half(self):
        if self._is(R_ZERO): return self
        return E(self.t, self.t.new(HALF, self.i))
    def inv(self): return E(self.t, self.t.new(INV, self.i))


# ---- the tower and the Miller steps, generic over the element type (formulas of csrc/pairing.cuh) -----------------------------
def f2_add(a, b): return (a[0] + b[0], a[1] + b[1])
def f2_sub(a, b): return (a[0] - b[0], a[1] - b[1])
def f2_neg(a): return (-a[0], -a[1])
def f2_dbl(a): return (a[0] + a[0], a[1] + a[1])
def f2_triple(a): return f2_add(f2_dbl(a), a)
def f2_half(a): return (a[0].half(), a[1].half())
def f2_conj(a): return (a[0], -a[1])
def f2_mul_xi(a): return (a[0] - a[1], a[0] + a[1])
def f2_scale(a, k): return (a[0] * k, a[1] * k)
def f2_mul(a, b):
    t0, t1, t2 = a[0] * b[0], a[1] * b[1], (a[0] + a[1]) * (b[0] + b[1])
    return (t0 - t1, t2 - t0 - t1)
def f2_sqr(a):
    t = a[0] * a[1]
    return ((a[0] + a[1]) * (a[0] - a[1]), t + t)
def f2_mul_b(a):                                     # * 4 (u + 1), the twist constant, without a multiplication
    x = f2_mul_xi(a)
    return f2_dbl(f2_dbl(x))
def f2_inv(a):
    d = (a[0] * a[0] + a[1] * a[1]).inv()
    return (a[0] * d, -(a[1] * d))

def f6_add(a, b): return tuple(f2_add(x, y) for x, y in zip(a, b))
def f6_sub(a, b): return tuple(f2_sub(x, y) for x, y in zip(a, b))
def f6_neg(a): return tuple(f2_neg(x) for x in a)
def f6_mul_v(a): return (f2_mul_xi(a[2]), a[0], a[1])
def f6_mul(a, b):
    v0, v1, v2 = f2_mul(a[0], b[0]), f2_mul(a[1], b[1]), f2_mul(a[2], b[2])
    t0 = f2_sub(f2_sub(f2_mul(f2_add(a[1], a[2]), f2_add(b[1], b[2])), v1), v2)
    t1 = f2_sub(f2_sub(f2_mul(f2_add(a[0], a[1]), f2_add(b[0], b[1])), v0), v1)
    t2 = f2_sub(f2_sub(f2_mul(f2_add(a[0], a[2]), f2_add(b[0], b[2])), v0), v2)
    return (f2_add(v0, f2_mul_xi(t0)), f2_add(t1, f2_mul_xi(v2)), f2_add(t2, v1))
def f6_inv(a):
    t0 = f2_sub(f2_sqr(a[0]), f2_mul_xi(f2_mul(a[1], a[2])))
    t1 = f2_sub(f2_mul_xi(f2_sqr(a[2])), f2_mul(a[0], a[1]))
    t2 = f2_sub(f2_sqr(a[1]), f2_mul(a[0], a[2]))
    d = f2_inv(f2_add(f2_mul(a[0], t0), f2_mul_xi(f2_add(f2_mul(a[2], t1), f2_mul(a[1], t2)))))
    return (f2_mul(t0, d), f2_mul(t1, d), f2_mul(t2, d))
def f6_mul_by_01(s, c0, c1):
    a_a, b_b = f2_mul(s[0], c0), f2_mul(s[1], c1)
    t1 = f2_add(f2_mul_xi(f2_sub(f2_mul(f2_add(s[1], s[2]), c1), b_b)), a_a)
    t3 = f2_add(f2_sub(f2_mul(f2_add(s[0], s[2]), c0), a_a), b_b)
    t2 = f2_sub(f2_sub(f2_mul(f2_add(s[0], s[1]), f2_add(c0, c1)), a_a), b_b)
    return (t1, t2, t3)
def f6_mul_by_1(s, c1): return (f2_mul_xi(f2_mul(s[2], c1)), f2_mul(s[0], c1), f2_mul(s[1], c1))

def f12_mul(a, b):
    aa, bb = f6_mul(a[0], b[0]), f6_mul(a[1], b[1])
    c1 = f6_sub(f6_sub(f6_mul(f6_add(a[0], a[1]), f6_add(b[0], b[1])), aa), bb)
    return (f6_add(aa, f6_mul_v(bb)), c1)
def f12_sqr(a):
    ab = f6_mul(a[0], a[1])
    t = f6_mul(f6_add(a[0], a[1]), f6_add(a[0], f6_mul_v(a[1])))
    return (f6_sub(f6_sub(t, ab), f6_mul_v(ab)), f6_add(ab, ab))
def f12_conj(a): return (a[0], f6_neg(a[1]))
def f12_inv(a):
    d = f6_inv(f6_sub(f6_mul(a[0], a[0]), f6_mul_v(f6_mul(a[1], a[1]))))
    return (f6_mul(a[0], d), f6_neg(f6_mul(a[1], d)))
def f12_mul_by_014(f, c0, c1, c4):
    aa, bb = f6_mul_by_01(f[0], c0, c1), f6_mul_by_1(f[1], c4)
    n1 = f6_sub(f6_sub(f6_mul_by_01(f6_add(f[1], f[0]), c0, f2_add(c1, c4)), aa), bb)
    return (f6_add(f6_mul_v(bb), aa), n1)
def fp4_sqr(x, y):
    t0, t1 = f2_sqr(x), f2_sqr(y)
    return f2_add(f2_mul_xi(t1), t0), f2_sub(f2_sub(f2_sqr(f2_add(x, y)), t0), t1)
def f12_cyc_sqr(a):
    z0, z4, z3 = a[0]; z2, z1, z5 = a[1]
    t0, t1 = fp4_sqr(z0, z1)
    z0 = f2_add(f2_dbl(f2_sub(t0, z0)), t0)
    z1 = f2_add(f2_dbl(f2_add(t1, z1)), t1)
    t0, t1 = fp4_sqr(z2, z3)
    t2, t3 = fp4_sqr(z4, z5)
    z4 = f2_add(f2_dbl(f2_sub(t0, z4)), t0)
    z5 = f2_add(f2_dbl(f2_add(t1, z5)), t1)
    t0 = f2_mul_xi(t3)
    z2 = f2_add(f2_dbl(f2_add(t0, z2)), t0)
    z3 = f2_add(f2_dbl(f2_sub(t2, z3)), t2)
    return ((z0, z4, z3), (z2, z1, z5))
def f6_frob(a, k, C):
    fr = (lambda x: f2_conj(x)) if k & 1 else (lambda x: x)
    return (fr(a[0]), f2_mul(fr(a[1]), C["G6_1_%d" % k]), f2_mul(fr(a[2]), C["G6_2_%d" % k]))
def f12_frob(a, k, C):
    c1 = f6_frob(a[1], k, C)
    return (f6_frob(a[0], k, C), tuple(f2_mul(c, C["G12_%d" % k]) for c in c1))

def doubling_step(r):
    X, Y, Z = r
    a = f2_half(f2_mul(X, Y))
    b, c = f2_sqr(Y), f2_sqr(Z)
    e = f2_mul_b(f2_triple(c))
    f = f2_triple(e)
    g = f2_half(f2_add(b, f))
    h = f2_sub(f2_sqr(f2_add(Y, Z)), f2_add(b, c))
    i = f2_sub(e, b)
    j = f2_sqr(X)
    e2 = f2_sqr(e)
    X3 = f2_mul(a, f2_sub(b, f))
    Y3 = f2_sub(f2_sqr(g), f2_triple(e2))
    Z3 = f2_mul(b, h)
    return (X3, Y3, Z3), (i, f2_triple(j), f2_neg(h))
def addition_step(r, q):
    X, Y, Z = r
    theta = f2_sub(Y, f2_mul(q[1], Z))
    lam = f2_sub(X, f2_mul(q[0], Z))
    c, d = f2_sqr(theta), f2_sqr(lam)
    e, f, g = f2_mul(lam, d), f2_mul(Z, c), f2_mul(X, d)
    h = f2_sub(f2_add(e, f), f2_dbl(g))
    X3 = f2_mul(lam, h)
    Y3 = f2_sub(f2_mul(theta, f2_sub(g, h)), f2_mul(e, Y))
    Z3 = f2_mul(Z, e)
    j = f2_sub(f2_mul(theta, q[0]), f2_mul(lam, q[1]))
    return (X3, Y3, Z3), (j, f2_neg(theta), lam)
def line_at(co, p):
    """the line as a sparse F_q12 element (c0, c1 x_P, c4 y_P) at tower positions 0, 1, 4"""
    return (co[0], f2_scale(co[1], p[0]), f2_scale(co[2], p[1]))
def sparse_to_f12(l, zero):
    return ((l[0], l[1], zero), (zero, l[2], zero))
def mul_lines(l, m, zero):
    """(c0 + c1 v + c4 v w)(d0 + d1 v + d4 v w): five non-zero coefficients, six F_q2 products"""
    c0d0, c1d1, c4d4 = f2_mul(l[0], m[0]), f2_mul(l[1], m[1]), f2_mul(l[2], m[2])
    v1 = f2_sub(f2_sub(f2_mul(f2_add(l[0], l[1]), f2_add(m[0], m[1])), c0d0), c1d1)          # c0 d1 + c1 d0     (v)
    vw = f2_sub(f2_sub(f2_mul(f2_add(l[0], l[2]), f2_add(m[0], m[2])), c0d0), c4d4)          # c0 d4 + c4 d0     (v w)
    v2w = f2_sub(f2_sub(f2_mul(f2_add(l[1], l[2]), f2_add(m[1], m[2])), c1d1), c4d4)         # c1 d4 + c4 d1     (v^2 w)
    return ((f2_add(c0d0, f2_mul_xi(c4d4)), v1, c1d1), (zero, vw, v2w))                       # c4 d4 v^2 w^2 = xi c4 d4


# ---- segments -----------------------------------------------------------------------------------------------------------
def blk12(t, regs):
    e = [E(t, t.pin(r)) for r in regs]
    return (((e[0], e[1]), (e[2], e[3]), (e[4], e[5])), ((e[6], e[7]), (e[8], e[9]), (e[10], e[11])))
def flat12(a): return [c for f6 in a for f2 in f6 for c in f2]
def blk_g2proj(t, regs):
    e = [E(t, t.pin(r)) for r in regs]
    return ((e[0], e[1]), (e[2], e[3]), (e[4], e[5]))
def consts(t): return {nm: (E(t, t.pin(rr[0])), E(t, t.pin(rr[1]))) for nm, (rr, _) in CONST_F2.items()}

def seg_miller(t, bits, first):
    """len(bits) consecutive iterations of the Miller loop for NPAIRS pairs (bits[i] = 1: iteration i also adds Q_k):
    f <- f^2 * prod lines,  R_k <- 2 R_k (+ Q_k).  Several iterations per segment let the scheduler run the chain of the points
    R_k (two levels of multiplications per iteration, independent of f) AHEAD of the chain of f, so that a segment of n iterations
    needs ~2n + 3 levels of multiplications instead of 5n."""
    zero2 = (E(t, t.pin(R_ZERO)), E(t, t.pin(R_ZERO)))
    one = E(t, t.pin(R_ONE))
    f = (((one, zero2[0]), zero2, zero2), (zero2, zero2, zero2)) if first else blk12(t, BLOCKS["F"])
    ps, qs, rs = [], [], []
    for k in range(NPAIRS):
        ps.append((E(t, t.pin(R_P[k][0])), E(t, t.pin(R_P[k][1]))))
        q = ((E(t, t.pin(R_Q[k][0])), E(t, t.pin(R_Q[k][1]))), (E(t, t.pin(R_Q[k][2])), E(t, t.pin(R_Q[k][3]))))
        qs.append(q)
        rs.append((q[0], q[1], (one, zero2[0])) if first else blk_g2proj(t, R_R[k]))
    for with_add in bits:
        lines = []
        for k in range(NPAIRS):
            rs[k], co = doubling_step(rs[k])
            lines.append(line_at(co, ps[k]))
            if with_add:
                rs[k], co = addition_step(rs[k], qs[k])
                lines.append(line_at(co, ps[k]))
        f = f12_sqr(f)
        # lines are multiplied pairwise first (off the critical path of f), then folded into f
        prods = []
        for i in range(0, len(lines) - 1, 2):
            prods.append(mul_lines(lines[i], lines[i + 1], zero2))
        if len(lines) & 1:
            prods.append(sparse_to_f12(lines[-1], zero2))
        while len(prods) > 1:
            prods = [f12_mul(prods[i], prods[i + 1]) if i + 1 < len(prods) else prods[i] for i in range(0, len(prods), 2)]
        f = f12_mul(f, prods[0])
    outs = []
    for k in range(NPAIRS):
        outs += list(zip([c for f2 in rs[k] for c in f2], R_R[k]))
    outs += list(zip(flat12(f), BLOCKS["F"]))
    return outs

def seg_easy(t):
    """M = m = r^(q^2) r,  r = conj(f) / f  for  f = conj(F) (the Miller value for the negative curve parameter); X = B = m"""
    C = consts(t)
    f = f12_conj(blk12(t, BLOCKS["F"]))
    r = f12_mul(f12_conj(f), f12_inv(f))
    m = f12_mul(f12_frob(r, 2, C), r)
    fm = flat12(m)
    return list(zip(fm, BLOCKS["M"])) + list(zip(fm, BLOCKS["X"])) + list(zip(fm, BLOCKS["B"]))
def seg_cyc(t, times=1):
    """X <- X^(2^times) by Granger-Scott squarings (several per segment: the sums after one squaring merge with those before the next)"""
    x = blk12(t, BLOCKS["X"])
    for _ in range(times): x = f12_cyc_sqr(x)
    return list(zip(flat12(x), BLOCKS["X"]))
def seg_mulb(t): return list(zip(flat12(f12_mul(blk12(t, BLOCKS["X"]), blk12(t, BLOCKS["B"]))), BLOCKS["X"]))
def seg_glue(t, other, dst, how):
    """dst = conj(X) * g(other) with g = conj / frob1; X = B = dst   (conj(X) = other^x after the square-and-multiply loop)"""
    C = consts(t)
    o = blk12(t, BLOCKS[other])
    g = f12_conj(o) if how == "conj" else f12_frob(o, 1, C)
    v = flat12(f12_mul(f12_conj(blk12(t, BLOCKS["X"])), g))
    return list(zip(v, BLOCKS[dst])) + list(zip(v, BLOCKS["X"])) + list(zip(v, BLOCKS["B"]))
def seg_xx(t):
    v = flat12(f12_conj(blk12(t, BLOCKS["X"])))
    return list(zip(v, BLOCKS["X"])) + list(zip(v, BLOCKS["B"]))
def seg_last(t):
    """OUT = conj(X) frob2(c) conj(c) * m^2 m"""
    C = consts(t)
    c, m = blk12(t, BLOCKS["CV"]), blk12(t, BLOCKS["M"])
    d = f12_mul(f12_mul(f12_conj(blk12(t, BLOCKS["X"])), f12_frob(c, 2, C)), f12_conj(c))
    return list(zip(flat12(f12_mul(d, f12_mul(f12_cyc_sqr(m), m))), BLOCKS["OUT"]))


# ---- lowering: binary ADD/SUB trees -> n-ary signed sums (LINC) ------------------------------------------------------------
# A step of the interpreter costs a barrier, an operand fetch from shared memory and a modular correction whatever it computes, so
# chains of two-operand additions (a Karatsuba tower is ~13 of them deep per F_q12 product) are flattened: every value that a
# multiplication or an output needs becomes ONE signed sum of up to LINC_K registers over multiplication results / inputs,
# and only sums with more terms go through helper sums (sub-trees of the original expression, shared between outputs).
K_MUL, K_LINC, K_HALF, K_INV = 1, 2, 3, 4
LINC_K = int(os.environ.get("PAIRING_LINC_K", "8"))       # sources per sum; every source carries a multiplier 1..8
LINC_MAXW = 60                                               # bound on sum |coefficient| of one sum (range of the quotient estimate)
COST2 = {K_MUL: 10.0, K_LINC: 3.0, K_HALF: 1.0, K_INV: 400.0, IN: 0.0}

def lower(t, outs, name=""):
    sys.setrecursionlimit(100000)
    n = len(t.kind)
    is_lin = lambda i: t.kind[i] in (ADD, SUB)
    live = [False] * n
    stack = [e.i for e, _ in outs]
    while stack:
        i = stack.pop()
        if live[i]: continue
        live[i] = True
        for o in (t.a[i], t.b[i]):
            if o >= 0: stack.append(o)
    required = set()
    for i in range(n):
        if live[i] and t.kind[i] in (MUL, HALF, INV):
            for o in (t.a[i], t.b[i]):
                if o >= 0 and is_lin(o): required.add(o)
    for e, _ in outs:
        if is_lin(e.i): required.add(e.i)
    emitted = {}                                      # LIN node -> {atom: coefficient}
    def expansion(v, helpers, top=True, memo=None):
        if memo is None: memo = {}
        def rec(x, root):
            if not is_lin(x) or (x in helpers and not root): return {x: 1}
            if x in memo and not root: return memo[x]
            ea, eb = rec(t.a[x], False), rec(t.b[x], False)
            sg = 1 if t.kind[x] == ADD else -1
            r = dict(ea)
            for k, c in eb.items():
                r[k] = r.get(k, 0) + sg * c
                if r[k] == 0: del r[k]
            if not root: memo[x] = r
            return r
        return rec(v, True)
    def slots(e):                                     # fields of the operation word; a sum that is too heavy counts as too long
        f = sum((abs(c) + 7) // 8 for c in e.values())
        return f if sum(abs(c) for c in e.values()) <= LINC_MAXW else 10 ** 6
    def descendants(v, helpers):
        seen, stack, out = set(), [t.a[v], t.b[v]], []
        while stack:
            x = stack.pop()
            if x in seen or not is_lin(x): continue
            seen.add(x); out.append(x)
            if x not in helpers: stack += [t.a[x], t.b[x]]
        return out
    fields = lambda c: (abs(c) + 7) // 8
    syn = {}                                          # synthetic helper sums (no node of the trace): id -> {atom: coefficient}
    depth = {}                                        # emitted sum -> number of sum levels below it
    def term_depth(e): return 1 + max([depth.get(a, 0) for a in e], default=0)
    for v in sorted(required):
        helpers = set()
        while True:
            e = expansion(v, helpers)
            if slots(e) <= LINC_K: break
            # a helper: a sub-expression of v whose expansion over plain values (no helper below it: the chain of sums stays two
            # deep) fits one operation; the one that saves most fields wins, sums that exist already are preferred
            best, best_gain = None, 1.0
            for w in descendants(v, helpers):
                if w in helpers: continue
                ew = emitted[w] if w in emitted else expansion(w, set())
                if w in emitted and depth[w] > 1: continue
                sw = slots(ew)
                if sw > LINC_K: continue
                gain = sw - 1 + (0.5 if w in emitted else 0.0)
                if gain > best_gain: best, best_gain = w, gain
            if best is None: break
            if best not in emitted:
                emitted[best] = expansion(best, set()); depth[best] = 1
            helpers.add(best)
        if slots(e) > LINC_K:
            # no sub-expression helps any more: cut the remaining terms into groups of LINC_K fields (synthetic partial sums)
            items = sorted(e.items(), key=lambda kv: depth.get(kv[0], 0))
            while slots(dict(items)) > LINC_K:
                grp, f, w = {}, 0, 0
                rest = []
                for a, c in items:
                    if f + fields(c) <= LINC_K and w + abs(c) <= LINC_MAXW and (a not in emitted or True):
                        grp[a] = c; f += fields(c); w += abs(c)
                    else: rest.append((a, c))
                assert len(grp) >= 2, "a sum of %s cannot be cut (one coefficient above LINC_MAXW?)" % name
                sid = n + len(syn)
                syn[sid] = grp; depth[sid] = term_depth(grp)
                items = rest + [(sid, 1)]
                items.sort(key=lambda kv: depth.get(kv[0], 0))
            e = dict(items)
        emitted[v] = e; depth[v] = term_depth(e)
    # the lowered graph: id -> (kind, [(source id, sign)...]); LINC sources repeat for coefficients of magnitude > 1
    g = {}
    for i in range(n):
        if not live[i]: continue
        k = t.kind[i]
        if k == IN: g[i] = (IN, [])
        elif k == MUL: g[i] = (K_MUL, [(t.a[i], 0), (t.b[i], 0)])
        elif k == HALF: g[i] = (K_HALF, [(t.a[i], 0)])
        elif k == INV: g[i] = (K_INV, [(t.a[i], 0)])
        elif i in emitted:
            srcs = []
            for a, c in sorted(emitted[i].items()):
                m = abs(c)
                while m > 0:
                    srcs.append((a, (0 if c > 0 else 1) | ((min(m, 8) - 1) << 1)))      # sign | (multiplier - 1) << 1
                    m -= min(m, 8)
            if not srcs: srcs = [(t.pin(R_ZERO), 0)]
            g[i] = (K_LINC, srcs)
    def linc_srcs(e):
        srcs = []
        for a, c in sorted(e.items()):
            m = abs(c)
            while m > 0:
                srcs.append((a, (0 if c > 0 else 1) | ((min(m, 8) - 1) << 1)))
                m -= min(m, 8)
        return srcs
    for sid, e in syn.items(): g[sid] = (K_LINC, linc_srcs(e))
    zero = t.pin(R_ZERO)
    if zero not in g: g[zero] = (IN, [])
    # drop what nothing reaches any more (LIN nodes that were inlined everywhere)
    keep, stack = set(), [e.i for e, _ in outs]
    while stack:
        i = stack.pop()
        if i in keep: continue
        keep.add(i)
        stack += [a for a, _ in g[i][1]]
    return {i: g[i] for i in keep}


# ---- scheduling + register allocation -----------------------------------------------------------------------------------
def compile_segment(name, build, nreg_cap=512):
    t = Trace()
    outs = build(t)                                   # [(E, pinned reg)]
    g = lower(t, outs, name)
    ids = sorted(g)
    users = {i: [] for i in ids}
    for i in ids:
        for a in {a for a, _ in g[i][1]}: users[a].append(i)
    prio = {}
    def get_prio(i):                                  # longest weighted path to an output (synthetic sums are not in id order)
        if i not in prio: prio[i] = COST2[g[i][0]] + max([get_prio(u) for u in users[i]], default=0.0)
        return prio[i]
    for i in ids: get_prio(i)
    step_of = {i: -1 for i in ids if g[i][0] == IN}
    unsched = {i for i in ids if g[i][0] != IN}
    steps = []
    while unsched:
        s = len(steps)
        ready = [i for i in unsched if all(a in step_of and step_of[a] < s for a, _ in g[i][1])]
        ready.sort(key=lambda i: (-prio[i], i))
        invs = [i for i in ready if g[i][0] == K_INV]
        muls = [i for i in ready if g[i][0] == K_MUL]
        lins = [i for i in ready if g[i][0] in (K_LINC, K_HALF)]
        if muls:
            # a step with multiplications costs several linear-only steps: while fewer than LANES multiplications are ready and a
            # ready sum is MORE urgent than every ready multiplication (it gates multiplications further up the critical path),
            # run the ready sums first; sums never share a step with multiplications (the warp would run both code paths in turn)
            if lins and len(muls) < LANES and (SCHED_LIN_FIRST == 1 or (SCHED_LIN_FIRST == 2 and prio[lins[0]] > prio[muls[0]])): chosen = lins[:LANES]
            else: chosen = muls[:LANES]
        elif lins: chosen = lins[:LANES]
        else: chosen = invs[:1]
        assert chosen, "scheduler stuck in " + name
        for i in chosen:
            step_of[i] = s; unsched.discard(i)
        steps.append(chosen)
    # register allocation
    last_use = {i: -1 for i in ids}
    for i in ids:
        for a, _ in g[i][1]: last_use[a] = max(last_use[a], step_of[i])
    out_regs = {}
    for e, r in outs: out_regs.setdefault(e.i, []).append(r)
    pin_readers = {t.reg[i]: [(step_of[u], u) for u in users[i]] for i in ids if g[i][0] == IN}
    written_pins = {}                                 # pinned reg -> step at which the segment overwrites it
    reg_of = {i: t.reg[i] for i in ids if g[i][0] == IN}
    free, next_tmp, peak, copies, release_at = [], [N_PINNED], [N_PINNED], [], {}
    def alloc_tmp():
        if free: return free.pop()
        r = next_tmp[0]; next_tmp[0] += 1; peak[0] = max(peak[0], next_tmp[0])
        return r
    enc = []
    for s, ops in enumerate(steps):
        for r in release_at.pop(s, []): free.append(r)         # registers whose last read was in an EARLIER step
        row = []
        for i in ops:
            dst = None
            pins = out_regs.get(i, [])
            for pr in pins:
                # direct write into a pinned output register: no OTHER lane may read its old value in this or a later step
                # (the operation itself may: a lane reads its operands before it writes)
                if dst is None and pr not in written_pins and all(us < s or (us == s and u == i) for us, u in pin_readers.get(pr, [])):
                    dst = pr; written_pins[pr] = s
                else:
                    copies.append((pr, i))
            if dst is None:
                dst = alloc_tmp()
                if not pins: release_at.setdefault(last_use[i] + 1, []).append(dst)
            reg_of[i] = dst
            row.append((g[i][0], dst, [(reg_of[a], sg) for a, sg in g[i][1]]))
        enc.append(row)
    for e, r in outs:
        if g[e.i][0] == IN and t.reg[e.i] != r:
            assert t.reg[e.i] not in written_pins, "pass-through source overwritten in " + name
            copies.append((r, e.i))
    copies = [(d, reg_of[sn]) for d, sn in copies if d != reg_of[sn]]
    # copies run after everything else, ordered so that every source is read before a copy overwrites it
    while copies:
        srcs_all = {s_ for _, s_ in copies}
        now = [(d, s_) for d, s_ in copies if d not in srcs_all][:LANES]
        if not now: raise RuntimeError("copy cycle in " + name)
        copies = [c for c in copies if c not in now]
        enc.append([(K_LINC, d, [(s_, 0)]) for d, s_ in now])
    assert peak[0] <= nreg_cap, (name, peak[0])
    return {"name": name, "steps": enc, "peak": peak[0], "mul_ops": sum(1 for row in enc for o in row if o[0] == K_MUL),
            "mul_steps": sum(1 for row in enc if any(o[0] == K_MUL for o in row)), "ops": sum(len(r) for r in enc)}


Q_TOP = Q >> 352
Q_RECIP = (1 << 44) // (((Q_TOP + 1) >> 10) + 1)       # k = ((T >> 10) * Q_RECIP) >> 44 <= floor(T / (Q_TOP + 1))
QOFF = 64 * Q                                           # start value of every signed sum: > LINC_MAXW q, so the total is positive

def linc_model(vals, flags):
    """the device's signed sum, limb for limb (csrc/pairing_coop.cuh pairing_linc); flags = sign | (multiplier - 1) << 1"""
    M = 0xFFFFFFFF
    acc = [(QOFF >> (32 * i)) & M for i in range(12)]
    top = QOFF >> 384
    wneg = 0
    for x, fl in zip(vals, flags):
        sg, mult = fl & 1, (fl >> 1) + 1
        m = M if sg else 0
        for i in range(12): acc[i] += (((x >> (32 * i)) & M) ^ m) * mult
        wneg += sg * mult
    fix = wneg * M
    T = (acc[11] - fix) + ((acc[10] - fix) >> 32) + (top << 32)
    assert 0 <= T < (1 << 40)
    k = ((T >> 10) * Q_RECIP) >> 44
    c, out = 0, 0
    for i in range(12):
        v = acc[i] - fix - k * ((Q >> (32 * i)) & M) + c
        out |= (v & M) << (32 * i); c = v >> 32
    assert c + top == 0, "signed sum: quotient estimate too small"
    exact = QOFF + sum((-x if fl & 1 else x) * ((fl >> 1) + 1) for x, fl in zip(vals, flags))
    assert out == exact - k * Q and 0 <= out < 3 * Q
    for _ in range(2):
        if out >= Q: out -= Q
    assert out < Q
    return out

def encode(op):
    """128 bits: kind[0:3) dst[3:12) n[12:16) 8 x (reg | sign << 9 | (multiplier - 1) << 10) [16:120)"""
    k, d, srcs = op
    assert 0 <= d < 512 and len(srcs) <= 8
    w = k | (d << 3) | (len(srcs) << 12)
    for j, (r, fl) in enumerate(srcs):
        assert 0 <= r < 512 and 0 <= fl < 16
        w |= (r | (fl << 9)) << (16 + 13 * j)
    return [(w >> (32 * i)) & 0xFFFFFFFF for i in range(4)]


# ---- Python model of the interpreter (plain residues: the constants are not in Montgomery form here) -------------------------
def run_segment(seg, regs):
    for row in seg["steps"]:
        res = []
        for k, d, srcs in row:
            v = [regs[r] for r, _ in srcs]
            assert all(0 <= x < Q for x in v)
            if k == K_MUL: out = v[0] * v[1] % Q
            elif k == K_LINC: out = linc_model(v, [fl for _, fl in srcs])   # exactly what the device does
            elif k == K_HALF: out = v[0] * T.TWO_INV % Q
            elif k == K_INV: out = pow(v[0], -1, Q) if v[0] else 0
            else: raise ValueError(k)
            res.append((d, out))
        ds = [d for d, _ in res]
        assert len(set(ds)) == len(ds)
        for j, (k, d, srcs) in enumerate(row):                 # a lane may update a register in place; nobody else may read it in that step
            for j2, (k2, d2, srcs2) in enumerate(row):
                assert j == j2 or d not in [r for r, _ in srcs2], "a step writes a register another lane reads"
        for d, out in res: regs[d] = out

def to12(regs, blk):
    v = [regs[r] for r in blk]
    return (((v[0], v[1]), (v[2], v[3]), (v[4], v[5])), ((v[6], v[7]), (v[8], v[9]), (v[10], v[11])))

def build_all():
    segs = {}
    order = []
    def add(name, fn):
        segs[name] = compile_segment(name, fn); order.append(name)
    bits = [int(b) for b in bin(P.X_ABS)[3:]]           # 63 iterations; bit = 1: the iteration also adds Q
    macro = []
    for i0 in range(0, len(bits), MILLER_GROUP):
        grp = tuple(bits[i0:i0 + MILLER_GROUP])
        name = "miller_%s%s" % ("first_" if i0 == 0 else "", "".join(map(str, grp)))
        if name not in segs: add(name, lambda t, grp=grp, first=(i0 == 0): seg_miller(t, grp, first))
        macro.append(name)
    add("easy", seg_easy)
    for k in CYC_RUNS: add("cyc%d" % k, lambda t, k=k: seg_cyc(t, k))
    add("mulb", seg_mulb)
    add("glue_a", lambda t: seg_glue(t, "M", "A", "conj"))
    add("glue_b", lambda t: seg_glue(t, "A", "BV", "conj"))
    add("glue_c", lambda t: seg_glue(t, "BV", "CV", "frob1"))
    add("xx", seg_xx)
    add("last", seg_last)
    expx, run = [], 0
    def flush(run):
        while run:                                     # a run of squarings as the largest segments available
            k = max(c for c in CYC_RUNS if c <= run)
            expx.append("cyc%d" % k); run -= k
    for b in bits:
        run += 1
        if b:
            flush(run); run = 0
            expx.append("mulb")
    flush(run)
    macro += ["easy"] + expx + ["glue_a"] + expx + ["glue_b"] + expx + ["glue_c"] + expx + ["xx"] + expx + ["last"]
    return segs, order, macro


def model_run(segs, macro, pairs):
    regs = [0] * 512
    regs[R_ONE] = 1
    for nm, (rr, val) in CONST_F2.items(): regs[rr[0]], regs[rr[1]] = val
    for k, (p, q) in enumerate(pairs):
        regs[R_P[k][0]], regs[R_P[k][1]] = p
        regs[R_Q[k][0]], regs[R_Q[k][1]] = q[0]
        regs[R_Q[k][2]], regs[R_Q[k][3]] = q[1]
    for nm in macro: run_segment(segs[nm], regs)
    return to12(regs, BLOCKS["OUT"])


def self_check(segs, macro):
    rnd = random.Random(77)
    for trial in range(2):
        a, b = rnd.randrange(1, P.R), rnd.randrange(1, P.R)
        p1, q1 = P.g1_mul(a, P.G1_GEN), P.g2_mul(b, P.G2_GEN)
        p2, q2 = P.g1_mul(rnd.randrange(1, P.R), P.G1_GEN), P.g2_mul(rnd.randrange(1, P.R), P.G2_GEN)
        got = model_run(segs, macro, [(p1, q1), (p2, q2)])
        assert got == T.final_exp(T.multi_miller([(p1, q1), (p2, q2)])), "lane program != one-thread twin"
        if trial == 0:
            assert got == P.gt_cubed(P.f12_mul(P.pairing(p1, q1), P.pairing(p2, q2))), "lane program != oracle"
    s = rnd.randrange(1, P.R)
    one = model_run(segs, macro, [(P.g1_mul(s, P.G1_GEN), P.G2_GEN), (P.g1_neg(P.G1_GEN), P.g2_mul(s, P.G2_GEN))])
    assert one == P.F12_ONE, "e(sG1, G2) e(-G1, sG2) != 1"
    print("lane programs == twin == oracle (2 random products, 1 cancelling product)")


def main():
    segs, order, macro = build_all()
    tot_steps = sum(len(segs[nm]["steps"]) for nm in macro)
    tot_msteps = sum(segs[nm]["mul_steps"] for nm in macro)
    for nm in order:
        s = segs[nm]
        print("%-14s steps %3d (with MUL %3d)  ops %4d  MUL ops %4d  peak regs %3d" % (nm, len(s["steps"]), s["mul_steps"], s["ops"], s["mul_ops"], s["peak"]))
    print("macro: %d segment runs, %d steps (%d with multiplications), LANES = %d, LINC_K = %d; model time %.2f ms" % (
        len(macro), tot_steps, tot_msteps, LANES, LINC_K, (tot_msteps * 1.45 + (tot_steps - tot_msteps) * 0.45) * 1e-3))
    self_check(segs, macro)
    nreg = max(s["peak"] for s in segs.values())
    L = ["// generated by tools/gen_pairing_prog.py - do not edit", "#pragma once", '#include "../arith.cuh"', "namespace vrfs {",
         "// Lane programs of the BLS12-381 pairing product check (csrc/pairing_coop.cuh): step tables of LANES operations of 128 bits:",
         "//   kind [0:3)  dst [3:12)  n [12:16)  8 x (register | sign << 9 | (multiplier - 1) << 10) [16:120)",
         "// over a register file of F_q elements; kinds 0 NOP, 1 MUL (2 sources), 2 LINC (sum of n sources, each +-(1..8) x), 3 HALF, 4 INV.",
         "struct PairingProg {",
         "  static constexpr int LANES = %d, NPAIRS = %d, NREG = %d, NSEG = %d, NMACRO = %d;" % (LANES, NPAIRS, nreg, len(order), len(macro)),
         "  static constexpr int R_ZERO = %d, R_ONE = %d, R_CONST0 = %d, R_P0 = %d, R_Q0 = %d, R_OUT = %d;" % (
             R_ZERO, R_ONE, CONST_F2["G6_1_1"][0][0], R_P[0][0], R_Q[0][0], BLOCKS["OUT"][0]),
         "  // signed sums: start value 64 q (limbs qoff(i), top limb QOFF_TOP) and the reciprocal of q's top bits (see pairing_linc)",
         "  static constexpr uint32_t QOFF_TOP = %du, Q_RECIP = %du;" % (QOFF >> 384, Q_RECIP),
         "  static HD_INLINE uint32_t qoff(int i) { constexpr uint32_t t[12] = {%s}; return t[i]; }" % ", ".join("0x%08xu" % ((QOFF >> (32 * i)) & 0xFFFFFFFF) for i in range(12)),
         "};"]
    # steps are stored densely: PAIRING_STEP_OFF[s] .. PAIRING_STEP_OFF[s + 1] are the operations of step s (lane l takes the l-th)
    offs, step_off, tab = [], [0], []
    for nm in order:
        offs.append(len(step_off) - 1)
        for row in segs[nm]["steps"]:
            for o in row: tab += encode(o)
            step_off.append(len(tab) // 4)
    offs.append(len(step_off) - 1)
    L.append("// segments: " + ", ".join("%d %s" % (i, nm) for i, nm in enumerate(order)))
    L.append("VRFS_GLOBAL_TABLE uint32_t PAIRING_SEG_OFF[%d] = {%s};" % (len(offs), ", ".join(map(str, offs))))
    L.append("VRFS_GLOBAL_TABLE uint8_t PAIRING_MACRO[%d] = {%s};" % (len(macro), ", ".join(str(order.index(m)) for m in macro)))
    L.append("VRFS_GLOBAL_TABLE uint32_t PAIRING_STEP_OFF[%d] = {" % len(step_off))
    for i in range(0, len(step_off), 24):
        L.append("  " + ", ".join(map(str, step_off[i:i + 24])) + ",")
    L.append("};")
    L.append("alignas(16) VRFS_GLOBAL_TABLE uint32_t PAIRING_OPS[%d] = {" % len(tab))
    for i in range(0, len(tab), 16):
        L.append("  " + ", ".join("0x%08xu" % x for x in tab[i:i + 16]) + ",")
    L.append("};")
    L.append("// the constants of registers R_CONST0 .. (F_q2 values c0 | c1, in this order): " + ", ".join(CONST_F2))
    L.append("}  // namespace vrfs")
    text = "\n".join(L) + "\n"
    if "--check" in sys.argv:                       # tests: the committed tables are what this script emits
        if not os.path.exists(OUT) or open(OUT).read() != text:
            sys.exit("%s is stale: run tools/gen_pairing_prog.py" % OUT)
        print("tables up to date")
        return
    if "--dry" in sys.argv: return
    open(OUT, "w").write(text)
    print("wrote %s (%d steps, %d operations, <= %d lanes, %d registers)" % (OUT, len(step_off) - 1, len(tab) // 4, LANES, nreg))


if __name__ == "__main__":
    main()
