"""GPU: the CUDA kernels' Ed25519 and P-256 arithmetic (variable-base and fixed-base paths) and the device RFC 6979 DRBG against
libsodium (PyNaCl) and OpenSSL (`cryptography`), 1000 random cases each - independent of the oracle."""
import numpy as np
import pytest

import third_party_pins as T
from oracle import pyref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import ark_ec_vrfs_b200 as vrfs
    e = vrfs.Engine(0)
    yield e
    e.close()


def _gen(curve, n):
    return np.tile(T.xy64(curve.G[0], curve.G[1]), (n, 1))


def test_gpu_ed25519_against_libsodium(eng):
    nb = pytest.importorskip("nacl.bindings")
    import ark_ec_vrfs_b200 as vrfs
    n = T.check_ed25519(lambda kb: eng.output(vrfs.ED25519, kb, _gen(R.ED25519, len(kb))), lambda mb, pts: eng.output(vrfs.ED25519, mb, pts), n=1000)
    assert n == 1000
    # the fixed-base path (16-bit window tables): Secret::from_seed's public key
    seeds = [b"gpu-pin-seed-%d" % i for i in range(1000)]
    sk, pk = eng.secret_from_seed(vrfs.ED25519, seeds)
    for i in range(len(seeds)):
        assert T.ed_rfc8032_encode(pk[i]).tobytes() == nb.crypto_scalarmult_ed25519_base_noclamp(sk[i].tobytes())


def test_gpu_p256_and_rfc6979_against_openssl(eng):
    pytest.importorskip("cryptography")
    from cryptography.hazmat.primitives.asymmetric import ec
    import ark_ec_vrfs_b200 as vrfs
    n, drbg = T.check_p256(lambda kb: eng.output(vrfs.P256, kb, _gen(R.P256, len(kb))), lambda mb, pts: eng.output(vrfs.P256, mb, pts),
                           lambda kb, pts: eng.nonce(vrfs.P256, kb, pts), lambda pts: eng.point_encode(vrfs.P256, pts), n=1000)
    assert n == 1000 and drbg == 500
    seeds = [b"gpu-pin-seed-%d" % i for i in range(1000)]
    sk, pk = eng.secret_from_seed(vrfs.P256, seeds)
    for i in range(0, len(seeds), 4):
        nums = ec.derive_private_key(int.from_bytes(sk[i].tobytes(), "little"), ec.SECP256R1()).public_key().public_numbers()
        assert T.xy64(nums.x, nums.y).tobytes() == pk[i].tobytes()
