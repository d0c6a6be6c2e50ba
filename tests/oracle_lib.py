"""ctypes binding of the CPU oracle (oracle/liboracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_lib = None
NTHREADS = max(1, os.cpu_count() or 1)

BANDERSNATCH, ED25519, P256, BANDERSNATCH_SW, JUBJUB, BABYJUBJUB = 0, 1, 2, 3, 4, 5


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "vrf_oracle.c")
        if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)):
            build()
        _lib = C.CDLL(so)
        for f in ("oracle_hash_len", "oracle_point_enc_len", "oracle_challenge_len", "oracle_ietf_signature_len", "oracle_pedersen_signature_len"):
            getattr(_lib, f).restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(x, shape=None):
    a = np.ascontiguousarray(np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray)) else x, dtype=np.uint8)
    return a if shape is None else a.reshape(shape)


def pack_var(items):
    """list of bytes -> (concatenated uint8 array, uint64 offsets[n+1])"""
    off = np.zeros(len(items) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(b) for b in items], dtype=np.uint64)
    data = np.frombuffer(b"".join(items), dtype=np.uint8).copy() if off[-1] else np.zeros(1, dtype=np.uint8)
    return data, off


def sha512(msg: bytes) -> bytes:
    out = np.zeros(64, np.uint8); m = _u8(msg) if msg else np.zeros(1, np.uint8)
    lib().oracle_sha512(_p(m), C.c_size_t(len(msg)), _p(out)); return out.tobytes()


def sha256(msg: bytes) -> bytes:
    out = np.zeros(32, np.uint8); m = _u8(msg) if msg else np.zeros(1, np.uint8)
    lib().oracle_sha256(_p(m), C.c_size_t(len(msg)), _p(out)); return out.tobytes()


def hmac_sha256(key: bytes, msg: bytes) -> bytes:
    out = np.zeros(32, np.uint8); k = _u8(key) if key else np.zeros(1, np.uint8); m = _u8(msg) if msg else np.zeros(1, np.uint8)
    lib().oracle_hmac_sha256(_p(k), C.c_size_t(len(key)), _p(m), C.c_size_t(len(msg)), _p(out)); return out.tobytes()


def secret_from_seed(suite, seeds, nthreads=NTHREADS):
    data, off = pack_var(seeds); n = len(seeds)
    sk = np.zeros((n, 32), np.uint8); pk = np.zeros((n, 64), np.uint8)
    lib().oracle_secret_from_seed_batch(suite, C.c_size_t(n), _p(data), _p(off), _p(sk), _p(pk), nthreads)
    return sk, pk


def point_encode(suite, pts, nthreads=NTHREADS):
    pts = _u8(pts, (-1, 64)); n = len(pts); L = lib().oracle_point_enc_len(suite)
    out = np.zeros((n, L), np.uint8)
    lib().oracle_point_encode_batch(suite, C.c_size_t(n), _p(pts), _p(out), nthreads); return out


def point_decode(suite, enc, nthreads=NTHREADS):
    L = lib().oracle_point_enc_len(suite); enc = _u8(enc, (-1, L)); n = len(enc)
    out = np.zeros((n, 64), np.uint8); ok = np.zeros(n, np.uint8)
    lib().oracle_point_decode_batch(suite, C.c_size_t(n), _p(enc), _p(out), _p(ok), nthreads); return out, ok


def data_to_point(suite, datas, nthreads=NTHREADS):
    data, off = pack_var(datas); n = len(datas)
    out = np.zeros((n, 64), np.uint8); ok = np.zeros(n, np.uint8)
    lib().oracle_data_to_point_batch(suite, C.c_size_t(n), _p(data), _p(off), _p(out), _p(ok), nthreads); return out, ok


def output(suite, sk, inp, nthreads=NTHREADS):
    sk = _u8(sk, (-1, 32)); inp = _u8(inp, (-1, 64)); n = len(sk); out = np.zeros((n, 64), np.uint8)
    lib().oracle_output_batch(suite, C.c_size_t(n), _p(sk), _p(inp), _p(out), nthreads); return out


def point_to_hash(suite, pts, nthreads=NTHREADS):
    pts = _u8(pts, (-1, 64)); n = len(pts); out = np.zeros((n, lib().oracle_hash_len(suite)), np.uint8)
    lib().oracle_point_to_hash_batch(suite, C.c_size_t(n), _p(pts), _p(out), nthreads); return out


def nonce(suite, sk, inp, nthreads=NTHREADS):
    sk = _u8(sk, (-1, 32)); inp = _u8(inp, (-1, 64)); n = len(sk); out = np.zeros((n, 32), np.uint8)
    lib().oracle_nonce_batch(suite, C.c_size_t(n), _p(sk), _p(inp), _p(out), nthreads); return out


def _ad(ads, n):
    if ads is None:
        return None, None
    if isinstance(ads, tuple):
        return ads
    assert len(ads) == n
    return pack_var(ads)


def ietf_prove(suite, sk, inp, outp, ads=None, nthreads=NTHREADS):
    sk = _u8(sk, (-1, 32)); inp = _u8(inp, (-1, 64)); outp = _u8(outp, (-1, 64)); n = len(sk)
    ad, off = _ad(ads, n); c = np.zeros((n, 32), np.uint8); s = np.zeros((n, 32), np.uint8)
    lib().oracle_ietf_prove_batch(suite, C.c_size_t(n), _p(sk), _p(inp), _p(outp), _p(ad), _p(off), _p(c), _p(s), nthreads)
    return c, s


def ietf_verify(suite, pk, inp, outp, c, s, ads=None, nthreads=NTHREADS, status=False):
    """status=True: (ok, per-item Result<(), Error>: 0 Ok / 1 VerificationFailure / 2 InvalidData)"""
    pk = _u8(pk, (-1, 64)); inp = _u8(inp, (-1, 64)); outp = _u8(outp, (-1, 64)); c = _u8(c, (-1, 32)); s = _u8(s, (-1, 32)); n = len(pk)
    ad, off = _ad(ads, n); ok = np.zeros(n, np.uint8); st = np.zeros(n, np.uint8) if status else None
    lib().oracle_ietf_verify_status_batch(suite, C.c_size_t(n), _p(pk), _p(inp), _p(outp), _p(c), _p(s), _p(ad), _p(off), _p(ok), _p(st), nthreads)
    return (ok, st) if status else ok


def pedersen_prove(suite, sk, inp, outp, ads=None, nthreads=NTHREADS):
    sk = _u8(sk, (-1, 32)); inp = _u8(inp, (-1, 64)); outp = _u8(outp, (-1, 64)); n = len(sk)
    ad, off = _ad(ads, n); proof = np.zeros((n, 256), np.uint8); bl = np.zeros((n, 32), np.uint8)
    lib().oracle_pedersen_prove_batch(suite, C.c_size_t(n), _p(sk), _p(inp), _p(outp), _p(ad), _p(off), _p(proof), _p(bl), nthreads)
    return proof, bl


def pedersen_verify(suite, inp, outp, proof, ads=None, nthreads=NTHREADS, status=False):
    inp = _u8(inp, (-1, 64)); outp = _u8(outp, (-1, 64)); proof = _u8(proof, (-1, 256)); n = len(inp)
    ad, off = _ad(ads, n); ok = np.zeros(n, np.uint8); st = np.zeros(n, np.uint8) if status else None
    lib().oracle_pedersen_verify_status_batch(suite, C.c_size_t(n), _p(inp), _p(outp), _p(proof), _p(ad), _p(off), _p(ok), _p(st), nthreads)
    return (ok, st) if status else ok


def msm_g1(bases, scalars, n_columns=1, nthreads=NTHREADS):
    bases = _u8(bases, (-1, 96)); n = len(bases); scalars = _u8(scalars, (n_columns * n, 32))
    out = np.zeros((n_columns, 96), np.uint8)
    lib().oracle_msm_g1(C.c_size_t(n), _p(bases), _p(scalars), n_columns, _p(out), nthreads); return out


def g1_mul_gen(scalars, nthreads=NTHREADS):
    scalars = _u8(scalars, (-1, 32)); n = len(scalars); out = np.zeros((n, 96), np.uint8)
    lib().oracle_g1_mul_gen_batch(C.c_size_t(n), _p(scalars), _p(out), nthreads); return out


# ---- wire formats ----------------------------------------------------------------------
def subgroup_check(suite, pts, nthreads=NTHREADS):
    pts = _u8(pts, (-1, 64)); n = len(pts); ok = np.zeros(n, np.uint8)
    lib().oracle_subgroup_check_batch(suite, C.c_size_t(n), _p(pts), _p(ok), nthreads); return ok


def point_decode_checked(suite, enc, nthreads=NTHREADS):
    L = lib().oracle_point_enc_len(suite); enc = _u8(enc, (-1, L)); n = len(enc)
    out = np.zeros((n, 64), np.uint8); ok = np.zeros(n, np.uint8)
    lib().oracle_point_decode_checked_batch(suite, C.c_size_t(n), _p(enc), _p(out), _p(ok), nthreads); return out, ok


def ietf_signature_len(suite):
    return int(lib().oracle_ietf_signature_len(suite))


def ietf_sign_wire(suite, sk, datas, ads=None, nthreads=NTHREADS):
    sk = _u8(sk, (-1, 32)); n = len(sk); data, off = pack_var(datas); ad, aoff = _ad(ads, n)
    sig = np.zeros((n, ietf_signature_len(suite)), np.uint8); ok = np.zeros(n, np.uint8)
    lib().oracle_ietf_sign_wire_batch(suite, C.c_size_t(n), _p(sk), _p(data), _p(off), _p(ad), _p(aoff), _p(sig), _p(ok), nthreads)
    return sig, ok


def ietf_verify_wire(suite, pk_enc, datas, sig, ads=None, want_hash=True, nthreads=NTHREADS, status=False):
    L = lib().oracle_point_enc_len(suite); pk_enc = _u8(pk_enc, (-1, L)); n = len(pk_enc)
    sig = _u8(sig, (n, ietf_signature_len(suite))); data, off = pack_var(datas); ad, aoff = _ad(ads, n)
    ok = np.zeros(n, np.uint8); h = np.zeros((n, lib().oracle_hash_len(suite)), np.uint8) if want_hash else None
    st = np.zeros(n, np.uint8) if status else None
    lib().oracle_ietf_verify_wire_status_batch(suite, C.c_size_t(n), _p(pk_enc), _p(data), _p(off), _p(sig), _p(ad), _p(aoff), _p(ok), _p(h), _p(st), nthreads)
    res = (ok, h) if want_hash else (ok,)
    res = res + (st,) if status else res
    return res if len(res) > 1 else res[0]


def pedersen_signature_len(suite):
    return int(lib().oracle_pedersen_signature_len(suite))


def pedersen_sign_wire(suite, sk, datas, ads=None, nthreads=NTHREADS):
    sk = _u8(sk, (-1, 32)); n = len(sk); data, off = pack_var(datas); ad, aoff = _ad(ads, n)
    sig = np.zeros((n, pedersen_signature_len(suite)), np.uint8); bl = np.zeros((n, 32), np.uint8); ok = np.zeros(n, np.uint8)
    lib().oracle_pedersen_sign_wire_batch(suite, C.c_size_t(n), _p(sk), _p(data), _p(off), _p(ad), _p(aoff), _p(sig), _p(bl), _p(ok), nthreads)
    return sig, bl, ok


def pedersen_verify_wire(suite, datas, sig, ads=None, nthreads=NTHREADS, status=False):
    sig = _u8(sig, (-1, pedersen_signature_len(suite))); n = len(sig); data, off = pack_var(datas); ad, aoff = _ad(ads, n)
    ok = np.zeros(n, np.uint8); st = np.zeros(n, np.uint8) if status else None
    lib().oracle_pedersen_verify_wire_status_batch(suite, C.c_size_t(n), _p(data), _p(off), _p(sig), _p(ad), _p(aoff), _p(ok), _p(st), nthreads)
    return (ok, st) if status else ok
