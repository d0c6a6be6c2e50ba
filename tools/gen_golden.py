#!/usr/bin/env python3
"""Writes tests/golden/*.json.

Two kinds of entries, kept apart by the "provenance" key:
  * "upstream"   - values recalled from the upstream crate's data/vectors/*.json (and RFC 9381
                   Appendix B) that an independent implementation reproduced bit-for-bit
                   (the fields marked with a tick).  These are literals in
                   this file and are NOT recomputed; they pin the oracle.
  * "regression" - values computed by oracle/pyref.py (deterministic); they pin the C oracle and
                   the CUDA engine to the Python model on suites / inputs with no upstream vector
                   (Ed25519; non-empty `ad`; TAI counters > 0).
Run from the repo root:  python tools/gen_golden.py
"""
import json, os, sys, hashlib
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import pyref as R

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")

UPSTREAM_BANDERSNATCH_IETF = [
    dict(comment="bandersnatch_sha-512_ell2_ietf - vector-1", seed="01", alpha="", salt="", ad="",
         sk="3d6406500d4009fdf2604546093665911e753f2213570a29521fd88bc30ede18",
         pk="a1b1da71cc4682e159b7da23050d8b6261eb11a3247c89b07ef56ccd002fd38b",
         h="c5eaf38334836d4b10e05d2c1021959a917e08eaf4eb46a8c4c8d1bec04e2c00",
         gamma="e7aa5154103450f0a0525a36a441f827296ee489ef30ed8787cff8df1bef223f",
         beta="fdeb377a4ffd7f95ebe48e5b43a88d069ce62188e49493500315ad55ee04d7442b93c4c91d5475370e9380496f4bc0b838c2483bce4e133c6f18b0adbb9e4722",
         proof_c="439fd9495643314fa623f2581f4b3d7d6037394468084f4ad7d8031479d9d101",
         proof_s="828bedd2ad95380b11f67a05ea0a76f0c3fef2bee9f043f4dffdddde09f55c01"),
    dict(comment="bandersnatch_sha-512_ell2_ietf - vector-2", seed="02", alpha="0a", salt="", ad="",
         sk="8b9063872331dda4c3c282f7d813fb3c13e7339b7dc9635fdc764e32cc57cb15",
         pk="5ebfe047f421e1a3e1d9bbb163839812657bbb3e4ffe9856a725b2b405844cf3",
         h="8c1d1425374f01d86b23bfeab770c60b58d2eeb9afc5900c8b8a918d09a6086b",
         gamma="60f32f5ad3e9694b82ccc0a735edb2f940f757ab333cc5f7b0a41158b80f574f",
         beta="44f3728bc5ad550aeeb89f8db340b2fceffc946be3e2d8c5d99b47c1fce344b3c7fcee223a9b29a64fe4a86a9994784bc165bb0fba03ca0a493f75bee89a0946"),
    dict(comment="bandersnatch_sha-512_ell2_ietf - vector-3", seed="03", alpha="", salt="", ad="0b8c",
         sk="6db187202f69e627e432296ae1d0f166ae6ac3c1222585b6ceae80ea07670b14",
         pk="9d97151298a5339866ddd3539d16696e19e6b68ac731562c807fe63a1ca49506",
         h="c5eaf38334836d4b10e05d2c1021959a917e08eaf4eb46a8c4c8d1bec04e2c00"),
]
UPSTREAM_BANDERSNATCH_PEDERSEN = [
    dict(comment="bandersnatch_sha-512_ell2_pedersen - vector-1", seed="01", alpha="", salt="", ad="",
         sk="3d6406500d4009fdf2604546093665911e753f2213570a29521fd88bc30ede18",
         h="c5eaf38334836d4b10e05d2c1021959a917e08eaf4eb46a8c4c8d1bec04e2c00",
         gamma="e7aa5154103450f0a0525a36a441f827296ee489ef30ed8787cff8df1bef223f",
         blinding="01371ac62e04d1faaadbebaa686aaf122143e2cda23aacbaa4796d206779a501",
         proof_pk_com="3b21abd58807bb6d93797001adaacd7113ec320dcf32d1226494e18a57931fc4",
         proof_r="8123054bfdb6918e0aa25c3337e6509eea262282fd26853bf7cd6db234583f5e",
         proof_ok="ac57ce6a53a887fc59b6aa73d8ff0e718b49bd9407a627ae0e9b9e7c5d0d175b",
         proof_s="0d379b65fb1e6b2adcbf80618c08e31fd526f06c2defa159158f5de146104c0f",
         proof_sb="e2ca83136143e0cac3f7ee863edd3879ed753b995b1ff8d58305d3b1f323630b"),
]
# RFC 9381 Appendix B.1 (ECVRF-P256-SHA256-TAI), Examples 10 and 11.  h2c data = enc(PK) || alpha.
UPSTREAM_P256 = [
    dict(comment="RFC 9381 Example 10", alpha=b"sample".hex(), ad="",
         sk="c9afa9d845ba75166b5c215767b1d6934e50c3db36e89b127b8a622b120f6721",
         pk="0360fed4ba255a9d31c961eb74c6356d68c049b8923b61fa6ce669622e60f29fb6",
         h="0272a877532e9ac193aff4401234266f59900a4a9e3fc3cfc6a4b7e467a15d06d4", tai_ctr=1,
         k="0d90591273453d2dc67312d39914e3a93e194ab47a58cd598886897076986f77",
         pi="035b5c726e8c0e2c488a107c600578ee75cb702343c153cb1eb8dec77f4b5071b4a53f0a46f018bc2c56e58d383f2305e0975972c26feea0eb122fe7893c15af376b33edf7de17c6ea056d4d82de6bc02f"),
    dict(comment="RFC 9381 Example 11", alpha=b"test".hex(), ad="",
         sk="c9afa9d845ba75166b5c215767b1d6934e50c3db36e89b127b8a622b120f6721",
         pk="0360fed4ba255a9d31c961eb74c6356d68c049b8923b61fa6ce669622e60f29fb6", tai_ctr=3,
         pi="034dac60aba508ba0c01aa9be80377ebd7562c4a52d74722e0abae7dc3080ddb56c19e067b15a8a8174905b13617804534214f935b94c2287f797e393eb0816969d864f37625b443f30f1a5a33f2b3c854",
         beta="a284f94ceec2ff4b3794629da7cbafa49121972671b466cab4ce170aa365f26d"),
]


def regression_vectors(S, n, with_pedersen=True):
    """Deterministic vectors from the Python model: seeds, alphas and ads of assorted lengths
    (ad lengths straddle the SHA-2 block boundaries)."""
    C = S.curve
    ad_lens = [0, 1, 2, 31, 32, 33, 50, 67, 68, 69, 111, 112, 127, 128, 129, 300]
    out = []
    for i in range(n):
        seed = hashlib.sha256(b"vrfs-b200-golden-seed" + bytes([i])).digest()[: 1 + i % 7]
        alpha = hashlib.sha512(b"vrfs-b200-golden-alpha" + bytes([i])).digest()[: (i * 5) % 64]
        adl = ad_lens[i % len(ad_lens)]
        ad = (hashlib.sha512(b"vrfs-b200-golden-ad" + bytes([i])).digest() * 5)[:adl]
        sk = R.secret_from_seed(S, seed)
        Y = C.mul(sk, C.G)
        salt = R.enc_pt(S, Y) if S.codec == "sec1" else b""
        if S.h2c == "tai":
            I, ctr = R.h2c_tai(S, salt + alpha, True)
        else:
            I, ctr = R.h2c_ell2(S, salt + alpha), None
        O = C.mul(sk, I)
        c, s = R.ietf_prove(S, sk, I, O, ad)
        assert R.ietf_verify(S, Y, I, O, ad, c, s)
        v = dict(comment=f"{S.name} regression {i}", seed=seed.hex(), alpha=alpha.hex(), salt=salt.hex(),
                 ad=ad.hex(), sk=R.enc_sc(S, sk).hex(), pk=R.enc_pt(S, Y).hex(), h=R.enc_pt(S, I).hex(),
                 gamma=R.enc_pt(S, O).hex(), beta=R.point_to_hash(S, O).hex(),
                 nonce=R.enc_sc(S, R.nonce(S, sk, I)).hex(),
                 proof_c=R.enc_sc(S, c).hex(), proof_s=R.enc_sc(S, s).hex())
        if ctr is not None:
            v["tai_ctr"] = ctr
        if with_pedersen:
            (Yb, Rr, Ok, ps, psb), b = R.pedersen_prove(S, sk, I, O, ad)
            assert R.pedersen_verify(S, I, O, ad, (Yb, Rr, Ok, ps, psb))
            v.update(blinding=R.enc_sc(S, b).hex(), ped_pk_com=R.enc_pt(S, Yb).hex(), ped_r=R.enc_pt(S, Rr).hex(),
                     ped_ok=R.enc_pt(S, Ok).hex(), ped_s=R.enc_sc(S, ps).hex(), ped_sb=R.enc_sc(S, psb).hex())
        out.append(v)
    return out


def msm_vectors():
    """Small BLS12-381 G1 MSMs computed by naive double-and-add (pure group law; unpinned vs upstream)."""
    C = R.BLS12_381_G1
    out = []
    for n in (1, 2, 7, 33):
        bases, scalars = [], []
        for j in range(n):
            t = int.from_bytes(hashlib.sha512(b"vrfs-b200-golden-msm-base" + j.to_bytes(8, "little")).digest(), "little") % C.r
            bases.append(C.mul(t or 1, C.G))
            scalars.append(int.from_bytes(hashlib.sha512(b"vrfs-b200-golden-msm-sc" + j.to_bytes(8, "little")).digest(), "little") % C.r)
        if n >= 7:
            scalars[1] = 0
            scalars[2] = 1
            scalars[3] = C.r - 1
            bases[5] = bases[4]            # repeated base
            bases[6] = C.neg(bases[4])     # and its negation
        res = R.msm(C, bases, scalars)
        out.append(dict(n=n, bases=[f"{P[0]:096x}{P[1]:096x}" for P in bases],
                        scalars=[f"{s:064x}" for s in scalars],
                        result=None if res is None else f"{res[0]:096x}{res[1]:096x}"))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    files = {
        "bandersnatch_upstream.json": dict(
            provenance="upstream", suite="bandersnatch",
            note="recalled from upstream data/vectors/bandersnatch_sha-512_ell2_{ietf,pedersen}.json and reproduced "
                 "independently (SURVEY.md Appendix B, ticked fields only)",
            ietf=UPSTREAM_BANDERSNATCH_IETF, pedersen=UPSTREAM_BANDERSNATCH_PEDERSEN),
        "p256_rfc9381.json": dict(provenance="upstream", suite="secp256r1",
                                  note="RFC 9381 Appendix B.1 Examples 10-11", ietf=UPSTREAM_P256),
        "bandersnatch_regression.json": dict(provenance="regression", suite="bandersnatch",
                                             vectors=regression_vectors(R.SUITE_BANDERSNATCH, 24)),
        "ed25519_regression.json": dict(provenance="regression", suite="ed25519",
                                        note="parity unpinned: no upstream Ed25519 vector is available offline",
                                        vectors=regression_vectors(R.SUITE_ED25519, 24)),
        "p256_regression.json": dict(provenance="regression", suite="secp256r1",
                                     vectors=regression_vectors(R.SUITE_P256, 24)),
        "msm_g1_regression.json": dict(provenance="regression", curve="bls12-381-g1",
                                       note="bases: x||y big-endian hex (96 B); scalars big-endian hex", cases=msm_vectors()),
    }
    for name, obj in files.items():
        with open(os.path.join(OUT, name), "w") as f:
            json.dump(obj, f, indent=1)
            f.write("\n")
        print("wrote", name)


if __name__ == "__main__":
    main()
