"""TEST INFRASTRUCTURE ONLY - big-integer Python model of the ark-ec-vrfs hot path.

Role: (1) executable specification that pins the hashing / encoding conventions
against the golden vectors of SURVEY.md Appendix B, (2) generator of the
fixtures under tests/golden/ and of the constant tables baked into the CUDA
code (tools/gen_consts.py), (3) slow cross-check of the C oracle
(oracle/vrf_oracle.c).  Only tests/, tools/ and oracle/ import this module;
nothing under ark_ec_vrfs_b200/ may.

Reference mapping.  The mounted reference (/root/reference) is a deprecation
stub whose only content is the re-export list at /root/reference/src/lib.rs:13-17
of `ark_vrf` 0.1.0 (/root/reference/Cargo.toml:12), which is NOT mounted.  Every
function below therefore cites (a) the name re-exported at src/lib.rs:13-17 that
it restates and (b) the paragraph of SURVEY.md Appendix A that fixes its
behaviour.  Parity status: Bandersnatch IETF + Pedersen and P-256 are pinned
by golden vectors (tests/golden/*.json); Ed25519 and the ring commitment are
"parity unpinned" (no upstream vector available offline).
"""
from __future__ import annotations

import hashlib
import hmac as _hmac
from dataclasses import dataclass
from typing import Optional, Tuple, List

Point = Optional[Tuple[int, int]]  # affine; None = point at infinity (SW only)


# --------------------------------------------------------------------------
# field helpers
# --------------------------------------------------------------------------
def inv_mod(x: int, p: int) -> int:
    return pow(x, -1, p)


def legendre(x: int, p: int) -> int:
    t = pow(x, (p - 1) // 2, p)
    return -1 if t == p - 1 else t


def sqrt_mod(n: int, p: int) -> Optional[int]:
    """Tonelli-Shanks (ark-ff `SqrtPrecomputation::TonelliShanks`); returns
    *a* root - callers normalise the sign, as the reference's callers do."""
    n %= p
    if n == 0:
        return 0
    if legendre(n, p) != 1:
        return None
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while legendre(z, p) != -1:
        z += 1
    m, c, t, r = s, pow(z, q, p), pow(n, q, p), pow(n, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c = i, b * b % p
        t, r = t * c % p, r * b % p
    return r


# --------------------------------------------------------------------------
# curves
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class TECurve:
    """a*x^2 + y^2 = 1 + d*x^2*y^2 over F_p; identity (0,1)."""
    name: str
    p: int
    a: int
    d: int
    r: int          # prime subgroup order
    h: int          # cofactor
    G: Tuple[int, int]
    kind: str = "te"

    def identity(self):
        return (0, 1)

    def is_identity(self, P):
        return P == (0, 1)

    def on_curve(self, P):
        x, y = P
        p = self.p
        return (self.a * x * x + y * y - 1 - self.d * x * x * y * y) % p == 0

    def add(self, P, Q):
        p = self.p
        x1, y1 = P
        x2, y2 = Q
        t = self.d * x1 * x2 * y1 * y2 % p
        return ((x1 * y2 + x2 * y1) * inv_mod(1 + t, p) % p,
                (y1 * y2 - self.a * x1 * x2) * inv_mod(1 - t, p) % p)

    def neg(self, P):
        return ((-P[0]) % self.p, P[1])

    def mul(self, k: int, P):
        R = (0, 1)
        while k:
            if k & 1:
                R = self.add(R, P)
            P = self.add(P, P)
            k >>= 1
        return R


@dataclass(frozen=True)
class SWCurve:
    """y^2 = x^3 + a*x + b over F_p; identity None."""
    name: str
    p: int
    a: int
    b: int
    r: int
    h: int
    G: Tuple[int, int]
    kind: str = "sw"

    def identity(self):
        return None

    def is_identity(self, P):
        return P is None

    def on_curve(self, P):
        if P is None:
            return True
        x, y = P
        return (y * y - (x * x * x + self.a * x + self.b)) % self.p == 0

    def add(self, P, Q):
        p = self.p
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if (P[1] + Q[1]) % p == 0:
                return None
            l = (3 * P[0] * P[0] + self.a) * inv_mod(2 * P[1], p) % p
        else:
            l = (Q[1] - P[1]) * inv_mod(Q[0] - P[0], p) % p
        x = (l * l - P[0] - Q[0]) % p
        return (x, (l * (P[0] - x) - P[1]) % p)

    def neg(self, P):
        return None if P is None else (P[0], (-P[1]) % self.p)

    def mul(self, k: int, P):
        R = None
        while k:
            if k & 1:
                R = self.add(R, P)
            P = self.add(P, P)
            k >>= 1
        return R


# SURVEY.md Appendix C (all [CHECKED])
BLS_FR = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
BANDERSNATCH = TECurve(
    "bandersnatch", BLS_FR, (-5) % BLS_FR,
    45022363124591815672509500913686876175488063829319466900776701791074614335719,
    0x1cfb69d4ca675f520cce760202687600ff8f87007419047174fd06b52876e7e1, 4,
    (0x29c132cc2c0b34c5743711777bbe42f32b79c022ad998465e1e71866a252ae18,
     0x2a6c669eda123e0f157d8b50badcd586358cad81eee464605e3167b6cc974166))
BANDERSNATCH_MONT_A = 29978822694968839326280996386011761570173833766074948509196803838190355340952
BANDERSNATCH_MONT_B = 25465760566081946422412445027709227188579564747101592991722834452325077642517
BANDERSNATCH_GLV_LAMBDA = 8913659658109529928382530854484400854125314752504019737736543920008458395397

P25519 = 2 ** 255 - 19
ED25519 = TECurve(
    "ed25519", P25519, P25519 - 1, (-121665 * inv_mod(121666, P25519)) % P25519,
    2 ** 252 + 27742317777372353535851937790883648493, 8,
    (15112221349535400772501151409588531511454012693041857206046113283949847762202,
     4 * inv_mod(5, P25519) % P25519))

P256_P = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff
P256 = SWCurve(
    "secp256r1", P256_P, P256_P - 3,
    0x5ac635d8aa3a93e7b3ebbd55769886bc651d06b0cc53b0f63bce3c3e27d2604b,
    0xffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551, 1,
    (0x6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296,
     0x4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5))

BLS_FQ = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
BLS12_381_G1 = SWCurve(
    "bls12-381-g1", BLS_FQ, 0, 4, BLS_FR, 0x396c8c005555e1568c00aaab0000aaab,
    (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
     0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1))


# --------------------------------------------------------------------------
# suites  (ark_vrf::Suite + ark_vrf::suites::*, named at src/lib.rs:13-17)
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class Suite:
    name: str
    suite_id: bytes
    clen: int
    hash_name: str           # "sha512" | "sha256"
    codec: str               # "ark" (LE, arkworks compressed) | "sec1" (BE, SEC1 compressed)
    h2c: str                 # "ell2" | "tai"
    nonce_kind: str          # "rfc8032" | "rfc6979"
    curve: object
    blinding_base: Tuple[int, int]
    h2c_id: bytes = b""

    def H(self, data: bytes) -> bytes:
        return hashlib.new(self.hash_name, data).digest()

    @property
    def pt_len(self) -> int:
        return 33 if (self.codec == "sec1" or self.curve.kind == "sw") else 32


SUITE_BANDERSNATCH = Suite(
    "bandersnatch", b"Bandersnatch_SHA-512_ELL2", 32, "sha512", "ark", "ell2", "rfc8032", BANDERSNATCH,
    (6150229251051246713677296363717454238956877613358614224171740096471278798312,
     28442734166467795856797249030329035618871580593056783094884474814923353898473),
    b"Bandersnatch_XMD:SHA-512_ELL2_RO_")
SUITE_ED25519 = Suite(
    "ed25519", b"Ed25519_SHA-512_TAI", 16, "sha512", "ark", "tai", "rfc8032", ED25519,
    (52417091031015867055192825304177001039906336859819158874861527659737645967040,
     24364467899048426341436922427697710961180476432856951893648702734568269272170))
SUITE_P256 = Suite(
    "secp256r1", b"\x01", 16, "sha256", "sec1", "tai", "rfc6979", P256,
    (55516455597544811540149985232155473070193196202193483189274003004283034832642,
     48580550536742846740990228707183741745344724157532839324866819111997786854582))


# ---- the remaining suites of `ark_vrf::suites` (bandersnatch_sw, jubjub, baby-jubjub).  Curve constants are those of
# ark-ed-on-bls12-381-bandersnatch (SWConfig), ark-ed-on-bls12-381 and ark-ed-on-bn254, recalled and CHECKED numerically (on the
# curve, generator of prime order r, cofactor).  The suite strings, CHALLENGE_LEN = 32, TAI hash-to-curve and the arkworks codec are
# [RECALL]; the Pedersen blinding bases of these suites are NOT recalled: the placeholders below are
# data_to_point(b"vrfs-b200 blinding base " || SUITE_ID) and must be replaced by the crate's constants.  PARITY UNPINNED.
BANDERSNATCH_SW = SWCurve(
    "bandersnatch_sw", BLS_FR,
    10773120815616481058602537765553212789256758185246796157495669123169359657269,
    29569587568322301171008055308580903175558631321415017492731745847794083609535,
    BANDERSNATCH.r, 4,
    (30900340493481298850216505686589334086208278925799850409469406976849338430199,
     12663882780877899054958035777720958383845500985908634476792678820121468453298))
JUBJUB = TECurve(
    "jubjub", BLS_FR, (-1) % BLS_FR, 19257038036680949359750312669786877991949435402254120286184196891950884077233,
    0x0e7db4ea6533afa906673b0101343b00a6682093ccc81082d0970e5ed6f72cb7, 8,
    (8076246640662884909881801758704306714034609987455869804520522091855516602923,
     13262374693698910701929044844600465831413122818447359594527400194675274060458))
BN254_FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
BABYJUBJUB = TECurve(
    "babyjubjub", BN254_FR, 1, 9706598848417545097372247223557719406784115219466060233080913168975159366771,
    2736030358979909402780800718157159386076813972158567259200215660948447373041, 8,
    (19698561148652590122159747500897617769866003486955115824547446575314762165298,
     19298250018296453272277890825869354524455968081175474282777126169995084727839))


def _placeholder_blinding_base(suite_id: bytes, curve):
    S0 = Suite("tmp", suite_id, 32, "sha512", "ark", "tai", "rfc8032", curve, (0, 0))
    return h2c_tai(S0, b"vrfs-b200 blinding base " + suite_id)


def _late_suites():
    out = {}
    for idx, (name, sid, curve) in {3: ("bandersnatch_sw", b"Bandersnatch_SW_SHA-512_TAI", BANDERSNATCH_SW),
                                    4: ("jubjub", b"JubJub_SHA-512_TAI", JUBJUB),
                                    5: ("babyjubjub", b"BabyJubJub_SHA-512_TAI", BABYJUBJUB)}.items():
        out[idx] = Suite(name, sid, 32, "sha512", "ark", "tai", "rfc8032", curve, _placeholder_blinding_base(sid, curve))
    return out


SUITES = {0: SUITE_BANDERSNATCH, 1: SUITE_ED25519, 2: SUITE_P256}


# --------------------------------------------------------------------------
# codecs  (ark_vrf::codec::{ArkworksCodec, Sec1Codec})
# --------------------------------------------------------------------------
def enc_sc(S: Suite, s: int) -> bytes:
    return s.to_bytes(32, "big" if S.codec == "sec1" else "little")


def dec_sc(S: Suite, b: bytes) -> int:
    """scalar_decode = from_{le,be}_bytes_mod_order."""
    return int.from_bytes(b, "big" if S.codec == "sec1" else "little") % S.curve.r


def enc_pt(S: Suite, P) -> bytes:
    if S.codec == "sec1":
        assert P is not None
        return bytes([2 + (P[1] & 1)]) + P[0].to_bytes(32, "big")
    if S.curve.kind == "sw":
        # arkworks short-Weierstrass compressed form [RECALL]: x little-endian over ceil((bits + 2 flag bits)/8) = 33 bytes, the top two
        # bits of the last byte are the flags: bit 7 = y is the larger root (y > (p-1)/2), bit 6 = point at infinity
        if P is None:
            return bytes(32) + b"\x40"
        return P[0].to_bytes(32, "little") + bytes([0x80 if P[1] > (S.curve.p - 1) // 2 else 0])
    p = S.curve.p
    e = bytearray(P[1].to_bytes(32, "little"))
    if P[0] > (p - 1) // 2:
        e[31] |= 0x80
    return bytes(e)


def dec_pt(S: Suite, b: bytes):
    """point_decode (compressed, on-curve checked, no subgroup check)."""
    C = S.curve
    p = C.p
    if S.codec == "sec1":
        if len(b) != 33 or b[0] not in (2, 3):
            return None
        x = int.from_bytes(b[1:], "big")
        if x >= p:
            return None
        y = sqrt_mod((x * x * x + C.a * x + C.b) % p, p)
        if y is None:
            return None
        if (y & 1) != (b[0] & 1):
            y = (p - y) % p
        return (x, y)
    if C.kind == "sw":
        if len(b) < 33:
            return None
        flags = b[32] >> 6
        if flags == 3:
            return None                               # both flags: SWFlags::from_u8 gives None
        x = int.from_bytes(b[:32], "little")
        if x >= p:
            return None
        if flags == 1:
            return None                               # the identity: no typed Public / Input / Output holds it
        y = sqrt_mod((x * x * x + C.a * x + C.b) % p, p)
        if y is None:
            return None
        if (y > (p - 1) // 2) != bool(flags & 2):
            y = (p - y) % p
        return (x, y)
    b = bytearray(b[:32])
    sign = b[31] >> 7
    b[31] &= 0x7F
    y = int.from_bytes(b, "little")
    if y >= p:
        return None
    den = (C.a - C.d * y * y) % p
    if den == 0:
        return None
    x2 = (1 - y * y) * inv_mod(den, p) % p
    x = sqrt_mod(x2, p)
    if x is None:
        return None
    if (x > (p - 1) // 2) != bool(sign):
        x = (-x) % p
    return (x, y)


# --------------------------------------------------------------------------
# hash-to-curve  (ark_vrf::utils::hash_to_curve_{ell2_rfc_9380,tai_rfc_9381})
# --------------------------------------------------------------------------
def expand_message_xmd_ark(S: Suite, msg: bytes, dst: bytes, n: int, zpad: int) -> bytes:
    """RFC 9380 5.3.1 with ark-ff's Z_pad length quirk (zpad = 48, not the hash block size)."""
    dstp = dst + bytes([len(dst)])
    b0 = S.H(bytes(zpad) + msg + n.to_bytes(2, "big") + b"\0" + dstp)
    bi = S.H(b0 + b"\1" + dstp)
    out = bi
    i = 2
    while len(out) < n:
        bi = S.H(bytes(x ^ y for x, y in zip(b0, bi)) + bytes([i]) + dstp)
        out += bi
        i += 1
    return out[:n]


def elligator2_bandersnatch(u: int):
    """ark-ec Elligator2Map::map_to_curve, TE output."""
    q = BLS_FR
    A, B, Z = BANDERSNATCH_MONT_A, BANDERSNATCH_MONT_B, 5
    k = B
    jk = A * inv_mod(B, q) % q
    ksqi = inv_mod(B * B % q, q)
    den = (1 + Z * u * u) % q
    x1 = (-jk) * inv_mod(den if den else 1, q) % q
    gx1 = (x1 ** 3 + jk * x1 * x1 + x1 * ksqi) % q
    x2 = (-x1 - jk) % q
    gx2 = (x2 ** 3 + jk * x2 * x2 + x2 * ksqi) % q
    if legendre(gx1, q) in (0, 1):
        x, y, sg = x1, sqrt_mod(gx1, q), 1
    else:
        x, y, sg = x2, sqrt_mod(gx2, q), 0
    if (y & 1) != sg:
        y = (-y) % q
    s, t = x * k % q, y * k % q
    tv1 = (s + 1) % q
    if tv1 * t % q == 0:
        return (0, 1)
    return (s * inv_mod(t, q) % q, (s - 1) * inv_mod(tv1, q) % q)


def h2c_ell2(S: Suite, data: bytes):
    assert S is SUITE_BANDERSNATCH
    q = S.curve.p
    dst = b"ECVRF_" + S.h2c_id + S.suite_id
    ub = expand_message_xmd_ark(S, data, dst, 96, 48)
    u0 = int.from_bytes(ub[:48], "big") % q
    u1 = int.from_bytes(ub[48:], "big") % q
    C = S.curve
    return C.mul(C.h, C.add(elligator2_bandersnatch(u0), elligator2_bandersnatch(u1)))


def h2c_tai(S: Suite, data: bytes, return_ctr: bool = False):
    C = S.curve
    for ctr in range(256):
        hs = S.H(S.suite_id + b"\x01" + data + bytes([ctr]) + b"\x00")
        P = dec_pt(S, b"\x02" + hs) if S.codec == "sec1" else dec_pt(S, hs[:S.pt_len])
        if P is None or P == "inf":
            continue
        P = C.mul(C.h, P)
        if C.is_identity(P):
            continue
        return (P, ctr) if return_ctr else P
    return (None, 256) if return_ctr else None


def data_to_point(S: Suite, data: bytes):
    """Suite::data_to_point / Input::new."""
    return h2c_ell2(S, data) if S.h2c == "ell2" else h2c_tai(S, data)


# --------------------------------------------------------------------------
# nonce / challenge / point_to_hash  (ark_vrf::utils::*)
# --------------------------------------------------------------------------
def nonce_rfc8032(S: Suite, sk: int, I) -> int:
    t = S.H(enc_sc(S, sk))[32:64]
    return int.from_bytes(S.H(t + enc_pt(S, I)), "little") % S.curve.r


def nonce_rfc6979(S: Suite, sk: int, I) -> int:
    n = S.curve.r
    h1 = S.H(enc_pt(S, I))
    V = b"\x01" * 32
    K = b"\x00" * 32
    xb = sk.to_bytes(32, "big")
    hb = (int.from_bytes(h1, "big") % n).to_bytes(32, "big")
    mac = lambda k, m: _hmac.new(k, m, S.hash_name).digest()
    K = mac(K, V + b"\x00" + xb + hb)
    V = mac(K, V)
    K = mac(K, V + b"\x01" + xb + hb)
    V = mac(K, V)
    while True:
        V = mac(K, V)
        k = int.from_bytes(V, "big")
        if 1 <= k < n:
            return k
        K = mac(K, V + b"\x00")
        V = mac(K, V)


def nonce(S: Suite, sk: int, I) -> int:
    return nonce_rfc8032(S, sk, I) if S.nonce_kind == "rfc8032" else nonce_rfc6979(S, sk, I)


def challenge(S: Suite, pts, ad: bytes) -> int:
    buf = S.suite_id + b"\x02" + b"".join(enc_pt(S, P) for P in pts) + ad + b"\x00"
    return int.from_bytes(S.H(buf)[:S.clen], "big") % S.curve.r


def point_to_hash(S: Suite, P) -> bytes:
    return S.H(S.suite_id + b"\x03" + enc_pt(S, P) + b"\x00")


def secret_from_seed(S: Suite, seed: bytes) -> int:
    """Secret::from_seed."""
    return int.from_bytes(S.H(seed), "little") % S.curve.r


# --------------------------------------------------------------------------
# IETF VRF  (ark_vrf::ietf::{Prover,Verifier})
# --------------------------------------------------------------------------
def ietf_prove(S: Suite, sk: int, I, O, ad: bytes):
    C = S.curve
    Y = C.mul(sk, C.G)
    k = nonce(S, sk, I)
    c = challenge(S, [Y, I, O, C.mul(k, C.G), C.mul(k, I)], ad)
    return c, (k + c * sk) % C.r


def ietf_verify(S: Suite, Y, I, O, ad: bytes, c: int, s: int) -> bool:
    C = S.curve
    U = C.add(C.mul(s, C.G), C.neg(C.mul(c, Y)))
    V = C.add(C.mul(s, I), C.neg(C.mul(c, O)))
    if U is None or V is None:  # SW identity cannot be encoded
        return False
    return challenge(S, [Y, I, O, U, V], ad) == c


# --------------------------------------------------------------------------
# Pedersen VRF  (ark_vrf::pedersen::{Prover,Verifier})
# --------------------------------------------------------------------------
def pedersen_blinding(S: Suite, sk: int, I, ad: bytes) -> int:
    buf = S.suite_id + b"\xCC" + enc_sc(S, sk) + enc_pt(S, I) + ad + b"\x00"
    return int.from_bytes(S.H(buf), "big") % S.curve.r


def pedersen_prove(S: Suite, sk: int, I, O, ad: bytes):
    C = S.curve
    B = S.blinding_base
    b = pedersen_blinding(S, sk, I, ad)
    k = nonce(S, sk, I)
    kb = nonce(S, b, I)
    Yb = C.add(C.mul(sk, C.G), C.mul(b, B))
    R = C.add(C.mul(k, C.G), C.mul(kb, B))
    Ok = C.mul(k, I)
    c = challenge(S, [Yb, I, O, R, Ok], ad)
    return (Yb, R, Ok, (k + c * sk) % C.r, (kb + c * b) % C.r), b


def pedersen_verify(S: Suite, I, O, ad: bytes, proof) -> bool:
    C = S.curve
    B = S.blinding_base
    Yb, R, Ok, s, sb = proof
    c = challenge(S, [Yb, I, O, R, Ok], ad)
    if C.add(Ok, C.mul(c, O)) != C.mul(s, I):
        return False
    return C.add(R, C.mul(c, Yb)) == C.add(C.mul(s, C.G), C.mul(sb, B))


# --------------------------------------------------------------------------
# ring commitment MSM  (ark_vrf::ring -> ark-ec VariableBaseMSM::msm)
# --------------------------------------------------------------------------
def msm(C, bases, scalars):
    acc = C.identity()
    for P, s in zip(bases, scalars):
        acc = C.add(acc, C.mul(s % C.r, P))
    return acc


# --------------------------------------------------------------------------
# derived constants used by the CUDA engine (not part of the reference's behaviour)
# --------------------------------------------------------------------------
def bandersnatch_endo_consts():
    """GLV endomorphism psi on the TE model: psi(x,y) = (c(1-y^2)/(x*y), b(y^2+b)/(y^2-b)),
    psi(P) = lambda*P on the prime-order subgroup.  b, c are solved from G and lambda*G."""
    C = BANDERSNATCH
    q = C.p
    P = C.G
    Q = C.mul(BANDERSNATCH_GLV_LAMBDA, P)
    y2 = P[1] * P[1] % q
    # yQ*(y2 - b) = b*(y2 + b)  ->  b^2 + b*(y2 + yQ) - yQ*y2 = 0
    disc = ((y2 + Q[1]) ** 2 + 4 * Q[1] * y2) % q
    sq = sqrt_mod(disc, q)
    P2 = C.mul(3, C.G)
    Q2 = C.mul(BANDERSNATCH_GLV_LAMBDA, P2)
    for root in (sq, q - sq):
        b = (-(y2 + Q[1]) + root) * inv_mod(2, q) % q
        yy = P2[1] * P2[1] % q
        if (yy - b) % q and b * (yy + b) % q * inv_mod((yy - b) % q, q) % q == Q2[1]:
            c = Q[0] * P[0] % q * P[1] % q * inv_mod((1 - y2) % q, q) % q
            assert c * (1 - yy) % q * inv_mod(P2[0] * P2[1] % q, q) % q == Q2[0]
            return b, c
    raise AssertionError("endomorphism constants not found")


def endo_bandersnatch(P):
    b, c = bandersnatch_endo_consts()
    q = BLS_FR
    x, y = P
    if x == 0:           # identity or the 2-torsion point (0,-1): both map to identity
        return (0, 1)
    y2 = y * y % q
    return (c * (1 - y2) % q * inv_mod(x * y % q, q) % q,
            b * (y2 + b) % q * inv_mod((y2 - b) % q, q) % q)


def glv_basis(r: int, lam: int):
    """Short lattice basis {(a1,b1),(a2,b2)} with a + b*lam = 0 (mod r), via the extended Euclid
    remainder sequence (Gallant-Lambert-Vanstone)."""
    rows = [(r, 0), (lam, 1)]  # (remainder, t) with remainder = s*r + t*lam
    while rows[-1][0] * rows[-1][0] >= r:
        qn = rows[-2][0] // rows[-1][0]
        rows.append((rows[-2][0] - qn * rows[-1][0], rows[-2][1] - qn * rows[-1][1]))
    r_l, t_l = rows[-2]
    r_l1, t_l1 = rows[-1]
    qn = r_l // r_l1
    r_l2, t_l2 = r_l - qn * r_l1, t_l - qn * t_l1
    v1 = (r_l1, -t_l1)
    cand = [(r_l, -t_l), (r_l2, -t_l2)]
    v2 = min(cand, key=lambda v: v[0] * v[0] + v[1] * v[1])
    for a, b in (v1, v2):
        assert (a + b * lam) % r == 0
    return v1, v2


def glv_decompose(k: int, r: int, lam: int, basis=None):
    (a1, b1), (a2, b2) = basis or glv_basis(r, lam)
    det = a1 * b2 - a2 * b1
    # round(k*b2/det), round(-k*b1/det)
    c1 = (2 * k * b2 + det) // (2 * det)
    c2 = (-2 * k * b1 + det) // (2 * det)
    k1 = k - c1 * a1 - c2 * a2
    k2 = -c1 * b1 - c2 * b2
    assert (k1 + k2 * lam - k) % r == 0
    return k1, k2


SUITES.update(_late_suites())
SUITE_BANDERSNATCH_SW, SUITE_JUBJUB, SUITE_BABYJUBJUB = SUITES[3], SUITES[4], SUITES[5]
