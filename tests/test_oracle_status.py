"""CPU: the oracle's statement of which `Error` variant the reference's verifiers return (`Result<_, Error>` with
`Error::{VerificationFailure, InvalidData}`), on inputs where the classification is unambiguous."""
import numpy as np
import pytest

import oracle_lib as O
import vectors as V


@pytest.mark.parametrize("suite", [O.BANDERSNATCH, O.ED25519, O.P256])
def test_ietf_verify_status_classes(suite):
    n = 12
    w = V.make_ietf_proofs(suite, n, "fixed32", corrupt=False)
    pk, inp, out, c, s = (w[k].copy() for k in ("pk", "inp", "out", "c", "s"))
    c[1, 0] ^= 1                      # a well-formed proof that does not check
    s[2, 3] ^= 8
    out[3] = out[4]                   # valid point, wrong output
    inp[5, 5] ^= 0x10                 # off the curve
    pk[6, :32] = 0xFF                 # non-canonical coordinate
    out[7] = 0                        # TE: (0,0) is no curve point; SW: the identity, which has no compressed encoding
    ok, st = O.ietf_verify(suite, pk, inp, out, c, s, w["ads"], status=True)
    assert st.tolist() == [0, 1, 1, 1, 0, 2, 2, 2, 0, 0, 0, 0]
    assert np.array_equal(ok == 1, st == 0)
    assert np.array_equal(ok, O.ietf_verify(suite, pk, inp, out, c, s, w["ads"]))


@pytest.mark.parametrize("suite", [O.BANDERSNATCH, O.P256])
def test_wire_status_classes(suite):
    n = 6
    sk, pk = O.secret_from_seed(suite, [b"st-%d" % i for i in range(n)])
    datas = [b"d%d" % i for i in range(n)]
    sig, ok = O.ietf_sign_wire(suite, sk, datas)
    assert ok.all()
    L = O.lib().oracle_point_enc_len(suite)
    sig = sig.copy()
    sig[1, L] ^= 1                    # challenge
    sig[2, -32:] = 0xFF               # s >= r: not a canonical scalar
    pk_enc = O.point_encode(suite, pk).copy()
    pk_enc[3, 0] = 0x07 if suite == O.P256 else pk_enc[3, 0]
    if suite != O.P256:
        pk_enc[3] = 0xFF              # y >= p
    okv, _, st = O.ietf_verify_wire(suite, pk_enc, datas, sig, status=True)
    assert st.tolist() == [0, 1, 2, 2, 0, 0] and np.array_equal(okv == 1, st == 0)
    psig, _, pok = O.pedersen_sign_wire(suite, sk, datas)
    psig = psig.copy(); psig[1, -5] ^= 1; psig[2, -64:-32] = 0xFF
    okp, stp = O.pedersen_verify_wire(suite, datas, psig, status=True)
    assert stp.tolist() == [0, 1, 2, 0, 0, 0] and np.array_equal(okp == 1, stp == 0)
