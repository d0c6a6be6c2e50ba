"""Independent implementations that happen to be in the image, used to pin group arithmetic and the RFC 6979 DRBG:
   PyNaCl 1.6 (libsodium):  crypto_scalarmult_ed25519_base_noclamp, crypto_scalarmult_ed25519_noclamp, crypto_core_ed25519_add
   cryptography 48 (OpenSSL): P-256 key derivation (k*G), ECDH (x of k*P), deterministic ECDSA (RFC 6979 nonces)
Test infrastructure only.  The same checks are applied to the CPU oracle (tests/test_third_party_pins.py, CPU) and to the CUDA
kernels (tests/test_gpu_third_party_pins.py)."""
import hashlib

import numpy as np

ED_P = 2 ** 255 - 19
ED_L = 2 ** 252 + 27742317777372353535851937790883648493
ED_D = (-121665 * pow(121666, -1, ED_P)) % ED_P
P256_N = 0xffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551


def ed_rfc8032_encode(pts):
    """(n, 64) affine x||y LE -> (n, 32) RFC 8032 encodings (y with the PARITY of x in bit 255 - libsodium's format, not arkworks')"""
    pts = np.asarray(pts, np.uint8).reshape(-1, 64)
    out = pts[:, 32:].copy()
    out[:, 31] |= (pts[:, 0] & 1) << 7
    return out


def ed_rfc8032_decode(enc):
    """32-byte RFC 8032 encoding -> (x, y) ints"""
    b = bytearray(bytes(enc)); sign = b[31] >> 7; b[31] &= 0x7F
    y = int.from_bytes(b, "little")
    u, v = (y * y - 1) % ED_P, (ED_D * y * y + 1) % ED_P
    x = (u * pow(v, 3, ED_P) * pow(u * pow(v, 7, ED_P) % ED_P, (ED_P - 5) // 8, ED_P)) % ED_P
    if (v * x * x - u) % ED_P != 0:
        x = x * pow(2, (ED_P - 1) // 4, ED_P) % ED_P
    assert (v * x * x - u) % ED_P == 0
    if (x & 1) != sign:
        x = ED_P - x
    return x, y


def xy64(x, y):
    return np.frombuffer(x.to_bytes(32, "little") + y.to_bytes(32, "little"), np.uint8)


def random_scalars(n, order, tag):
    vals = [int.from_bytes(hashlib.sha512(tag + i.to_bytes(4, "little")).digest(), "little") % order for i in range(n)]
    vals[0] = 1; vals[1] = order - 1; vals[2] = 2
    return vals, np.frombuffer(b"".join(v.to_bytes(32, "little") for v in vals), np.uint8).reshape(n, 32).copy()


def check_ed25519(base_mul, var_mul, n=1000):
    """base_mul(scalars (n,32)) -> (n,64) points k*G;  var_mul(scalars, points) -> (n,64) points k*P.  Both against libsodium."""
    import nacl.bindings as nb
    ks, kb = random_scalars(n, ED_L, b"pin-ed-k")
    pk = base_mul(kb)
    for i in range(n):
        want = nb.crypto_scalarmult_ed25519_base_noclamp(ks[i].to_bytes(32, "little"))
        assert ed_rfc8032_encode(pk[i]).tobytes() == want, ("k*G", i)
    # variable base: points P_i = k_i*G (in the prime-order subgroup, as libsodium requires), scalars m_i
    ms, mb = random_scalars(n, ED_L, b"pin-ed-m")
    out = var_mul(mb, pk)
    for i in range(n):
        want = nb.crypto_scalarmult_ed25519_noclamp(ms[i].to_bytes(32, "little"), ed_rfc8032_encode(pk[i]).tobytes())
        assert ed_rfc8032_encode(out[i]).tobytes() == want, ("m*P", i)
    # group law: (k + m)*G == k*G + m*G by libsodium's point addition
    sums = np.frombuffer(b"".join(((a + b) % ED_L).to_bytes(32, "little") for a, b in zip(ks, ms)), np.uint8).reshape(n, 32)
    ps = base_mul(sums); pm = base_mul(mb)
    for i in range(3, n, 7):
        if (ks[i] + ms[i]) % ED_L == 0:
            continue
        want = nb.crypto_core_ed25519_add(ed_rfc8032_encode(pk[i]).tobytes(), ed_rfc8032_encode(pm[i]).tobytes())
        assert ed_rfc8032_encode(ps[i]).tobytes() == want, ("add", i)
    return n


def check_p256(base_mul, var_mul, nonce, encode, n=1000):
    """P-256 against OpenSSL through `cryptography`: k*G, x(k*P) by ECDH, and the RFC 6979 nonce through deterministic ECDSA
    (the nonce of our suite is RFC 6979 with h1 = SHA-256(point_encode(I)); an ECDSA signature over the message point_encode(I)
    uses the same k, and its r is x(k*G) mod n)."""
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import ec
    from cryptography.hazmat.primitives.asymmetric.utils import decode_dss_signature
    ks, kb = random_scalars(n, P256_N, b"pin-p256-k")
    pk = base_mul(kb)
    keys = []
    for i in range(n):
        key = ec.derive_private_key(ks[i], ec.SECP256R1())
        nums = key.public_key().public_numbers()
        assert xy64(nums.x, nums.y).tobytes() == pk[i].tobytes(), ("k*G", i)
        keys.append(key)
    ms, mb = random_scalars(n, P256_N, b"pin-p256-m")
    out = var_mul(mb, pk)                                   # m_i * (k_i * G)
    for i in range(n):
        shared = ec.derive_private_key(ms[i], ec.SECP256R1()).exchange(ec.ECDH(), keys[i].public_key())
        assert out[i, :32].tobytes()[::-1] == shared, ("ECDH", i)
    # RFC 6979: nonce(sk = k_i, input point I_i = out_i) vs OpenSSL's deterministic ECDSA over the message point_encode(I_i)
    kn = nonce(kb, out)
    enc = encode(out)
    kg = base_mul(kn)
    checked = 0
    for i in range(0, n, 2):
        try:
            sig = keys[i].sign(enc[i].tobytes(), ec.ECDSA(hashes.SHA256(), deterministic_signing=True))
        except TypeError:                                   # an older `cryptography` without deterministic_signing
            return n, 0
        r, _ = decode_dss_signature(sig)
        x = int.from_bytes(kg[i, :32].tobytes(), "little")
        assert x % P256_N == r, ("rfc6979", i)
        checked += 1
    return n, checked
