"""CPU: the oracle's Ed25519 and P-256 group arithmetic and its RFC 6979 nonce against libsodium (PyNaCl) and OpenSSL
(`cryptography`) - implementations that are independent of this repository and of arkworks.  Pins what the upstream vectors do
not reach: the whole Ed25519 curve arithmetic (no upstream Ed25519 vector exists offline) and the HMAC-DRBG beyond the two
RFC 9381 examples."""
import numpy as np
import pytest

import oracle_lib as O
import third_party_pins as T
from oracle import pyref as R


def _gen(curve, n):
    return np.tile(T.xy64(curve.G[0], curve.G[1]), (n, 1))


def test_oracle_ed25519_against_libsodium():
    pytest.importorskip("nacl.bindings")
    n = T.check_ed25519(lambda kb: O.output(O.ED25519, kb, _gen(R.ED25519, len(kb))), lambda mb, pts: O.output(O.ED25519, mb, pts), n=1000)
    assert n == 1000


def test_oracle_p256_and_rfc6979_against_openssl():
    pytest.importorskip("cryptography")
    n, drbg = T.check_p256(lambda kb: O.output(O.P256, kb, _gen(R.P256, len(kb))), lambda mb, pts: O.output(O.P256, mb, pts),
                           lambda kb, pts: O.nonce(O.P256, kb, pts), lambda pts: O.point_encode(O.P256, pts), n=1000)
    assert n == 1000 and drbg == 500


def test_ed25519_secret_from_seed_public_key_is_sk_times_base():
    """Secret::from_seed -> public: pk = sk * G with sk = LE(SHA-512(seed)) mod L, the base multiplication by libsodium"""
    nb = pytest.importorskip("nacl.bindings")
    seeds = [b"pin-seed-%d" % i for i in range(200)]
    sk, pk = O.secret_from_seed(O.ED25519, seeds)
    for i in range(len(seeds)):
        assert T.ed_rfc8032_encode(pk[i]).tobytes() == nb.crypto_scalarmult_ed25519_base_noclamp(sk[i].tobytes())
