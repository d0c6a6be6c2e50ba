"""GPU tests of the drop-in boundary: the per-item `Error` variant next to every verdict, the serialised 160-byte
Pedersen proof, secret hygiene of the staging buffers, lazy per-suite tables, argument checking of the Python wrapper, and
the multi-GPU forms (one caller / several GPUs, and the device-side exchange of the per-rank MSM partials)."""
import numpy as np
import pytest

import oracle_lib as O
import vectors as V

SUITES = [O.BANDERSNATCH, O.ED25519, O.P256, O.BANDERSNATCH_SW, O.JUBJUB, O.BABYJUBJUB]


@pytest.fixture(scope="module")
def eng():
    import ark_ec_vrfs_b200 as vrfs
    e = vrfs.Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("suite", SUITES)
def test_ietf_verify_reports_the_error_variant(eng, suite):
    """Result<(), Error>: Ok / VerificationFailure / InvalidData per item, the same classification as the oracle"""
    n = 400
    w = V.make_ietf_proofs(suite, n, "ragged")                       # corrupts c, s, output, pk (valid points) and input (off-curve)
    pk, inp, out = w["pk"].copy(), w["inp"].copy(), w["out"].copy()
    pk[1] = 0                                                       # TE: (0, 0) is off the curve; SW: the un-encodable identity
    out[2, :32] = 0xFF                                              # non-canonical x
    ok_o, st_o = O.ietf_verify(suite, pk, inp, out, w["c"], w["s"], w["ads"], status=True)
    ok, st = eng.ietf_verify(suite, pk, inp, out, w["c"], w["s"], w["ads"], status=True)
    assert np.array_equal(ok, ok_o) and np.array_equal(st, st_o)
    assert set(st.tolist()) == {0, 1, 2} and st[1] == 2 and st[2] == 2
    assert np.array_equal(ok == 1, st == 0)
    # the status array is optional and does not change the verdicts
    assert np.array_equal(eng.ietf_verify(suite, pk, inp, out, w["c"], w["s"], w["ads"]), ok)


@pytest.mark.gpu
@pytest.mark.parametrize("suite", SUITES)
def test_pedersen_verify_reports_the_error_variant(eng, suite):
    n = 210
    sk, pk, inp, out = V.make_keys_inputs(suite, n)
    ads = V.make_ads(n, "fixed32")
    pr, _ = O.pedersen_prove(suite, sk, inp, out, ads)
    pr = pr.copy(); out = out.copy()
    for i in range(0, n, 3):
        kind = (i // 3) % 6
        if kind < 3: pr[i, 64 * kind + 5] ^= 1                      # pk_com / R / Ok off the curve -> InvalidData
        elif kind == 3: pr[i, 200] ^= 4                              # s -> VerificationFailure
        elif kind == 4: out[i] = out[(i + 1) % n]                    # wrong but valid output -> VerificationFailure
        else: pr[i, 128:192] = pr[(i + 1) % n, 128:192]              # wrong but valid Ok
    ok_o, st_o = O.pedersen_verify(suite, inp, out, pr, ads, status=True)
    ok, st = eng.pedersen_verify(suite, inp, out, pr, ads, status=True)
    assert np.array_equal(ok, ok_o) and np.array_equal(st, st_o) and set(st.tolist()) == {0, 1, 2}


@pytest.mark.gpu
@pytest.mark.parametrize("suite", SUITES)
def test_wire_verifiers_report_the_error_variant(eng, suite):
    n = 160
    seeds = [b"wire-st-%d" % i for i in range(n)]
    sk, pk = O.secret_from_seed(suite, seeds)
    datas = [b"in-%d" % i for i in range(n)]
    pk_enc = O.point_encode(suite, pk)
    sig, ok = O.ietf_sign_wire(suite, sk, datas)
    assert ok.all()
    L = O.lib().oracle_point_enc_len(suite)
    sig = sig.copy(); pk_enc = pk_enc.copy()
    for i in range(0, n, 4):
        kind = (i // 4) % 4
        if kind == 0: sig[i, L + 2] ^= 1                             # challenge -> VerificationFailure
        elif kind == 1: sig[i, -32:] = 0xFF                          # s >= r -> InvalidData
        elif kind == 2: sig[i, 1:L] = sig[i, 1:L] ^ 0x5A             # Output bytes: almost surely no valid encoding -> InvalidData
        else: pk_enc[i] = pk_enc[(i + 1) % n]                        # another valid key -> VerificationFailure
    ok_o, h_o, st_o = O.ietf_verify_wire(suite, pk_enc, datas, sig, status=True)
    ok_g, h_g, st_g = eng.ietf_verify_wire(suite, pk_enc, datas, sig, status=True)
    assert np.array_equal(ok_g, ok_o) and np.array_equal(h_g, h_o) and np.array_equal(st_g, st_o)
    assert set(st_g.tolist()) == {0, 1, 2}
    psig, _, pok = O.pedersen_sign_wire(suite, sk, datas)
    assert pok.all()
    psig = psig.copy()
    psig[0, -1] ^= 1 if suite != O.P256 else 0; psig[0, -40] ^= 2      # sb / s bytes
    psig[1, L + 3] ^= 1                                              # pk_com
    psig[2, -64:-32] = 0xFF                                          # s not canonical
    ok_o, st_o = O.pedersen_verify_wire(suite, datas, psig, status=True)
    ok_g, st_g = eng.pedersen_verify_wire(suite, datas, psig, status=True)
    assert np.array_equal(ok_g, ok_o) and np.array_equal(st_g, st_o) and st_g[2] == 2 and st_g[3] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("suite", SUITES)
def test_pedersen_serialised_proof_form(eng, suite):
    """the typed pedersen::Proof as its CanonicalSerialize bytes (160 B for Bandersnatch): equals the tail of the oracle's wire
    signature, verifies, and rejects like the wire verifier"""
    n = 120
    seeds = [b"p160-%d" % i for i in range(n)]
    sk, _ = O.secret_from_seed(suite, seeds)
    datas = [b"alpha-%d" % i for i in range(n)]
    ads = V.make_ads(n, "ragged")
    sig_o, bl_o, ok_o = O.pedersen_sign_wire(suite, sk, datas, ads)
    assert ok_o.all()
    L = O.lib().oracle_point_enc_len(suite)
    inp, _ = O.data_to_point(suite, datas)
    out = O.output(suite, sk, inp)
    proof, bl = eng.pedersen_prove_compressed(suite, sk, inp, out, ads)
    assert proof.shape == (n, eng.pedersen_proof_len(suite)) and eng.pedersen_proof_len(suite) == 3 * L + 64
    if suite == O.BANDERSNATCH:
        assert proof.shape[1] == 160
    assert np.array_equal(proof, sig_o[:, L:]) and np.array_equal(bl, bl_o)
    proof = proof.copy()
    proof[0, 3 * L + 1] ^= 1; proof[1, 2] ^= 1; proof[2, 3 * L:3 * L + 32] = 0xFF
    ok, st = eng.pedersen_verify_compressed(suite, inp, out, proof, ads, status=True)
    full = np.concatenate([O.point_encode(suite, out), proof], axis=1)
    ok_w, st_w = O.pedersen_verify_wire(suite, datas, full, ads, status=True)
    assert np.array_equal(ok, ok_w) and np.array_equal(st, st_w) and ok[3:].all() and not ok[:3].any()


@pytest.mark.gpu
def test_secret_staging_buffers_are_wiped(eng):
    """device buffers that held `sk` (and nonces, blinding factors) are zero once the call has returned; buffers that
    held public values are not (which shows the hook reads what it claims to read)"""
    import ark_ec_vrfs_b200 as vrfs
    n = 300
    sk, pk, inp, out = V.make_keys_inputs(O.BANDERSNATCH, n)
    eng.ietf_prove(vrfs.BANDERSNATCH, sk, inp, out)
    assert not eng.debug_read_staging(0, n * 32).any()               # BUF_IN0: sk
    assert np.array_equal(eng.debug_read_staging(1, n * 64).reshape(n, 64), inp)      # BUF_IN1: the (public) input points
    assert not eng.debug_read_staging(9, n * 32).any()               # BUF_W0: the nonces k
    eng.pedersen_prove(vrfs.BANDERSNATCH, sk, inp, out)
    assert not eng.debug_read_staging(0, n * 32).any() and not eng.debug_read_staging(9, n * 96).any()
    assert not eng.debug_read_staging(8, n * 32).any()               # BUF_OUT1: the blinding factors handed back to the prover
    eng.output(vrfs.BANDERSNATCH, sk, inp)
    assert not eng.debug_read_staging(0, n * 32).any()
    seeds = [b"seed-%d" % i for i in range(n)]
    eng.secret_from_seed(vrfs.BANDERSNATCH, seeds)
    assert not eng.debug_read_staging(7, n * 32).any()               # BUF_OUT0: the derived secrets


@pytest.mark.gpu
def test_piecewise_host_calls_equal_small_calls(eng):
    """Host-buffer calls above three resident waves are cut into pieces whose copies overlap the kernels (two pieces; six for
    Pedersen prove): every item's result must be the one the same item gets in a small single-piece call.  Inputs and outputs live
    in page-locked memory from the library's own allocator (vrfs_host_alloc), one input array is page-locked in place
    (vrfs_host_register)."""
    import ctypes as C
    import ark_ec_vrfs_b200 as vrfs
    from ark_ec_vrfs_b200 import _lib
    base, n, chunk = 1024, 250_000, 50_000
    sk0, pk0, inp0, out0 = V.make_keys_inputs(O.BANDERSNATCH, base)
    idx = np.arange(n) % base
    sk, pk, inp, out = (vrfs.host_copy(a[idx]) for a in (sk0, pk0, inp0, out0))
    ads = [int(i).to_bytes(8, "little") for i in range(n)]          # distinct transcripts, so distinct proofs
    c, s = vrfs.host_buffer((n, 32)), vrfs.host_buffer((n, 32))
    eng.ietf_prove(vrfs.BANDERSNATCH, sk, inp, out, ads, out=(c, s))
    proof, bl = vrfs.host_buffer((n, 256)), vrfs.host_buffer((n, 32))
    eng.pedersen_prove(vrfs.BANDERSNATCH, sk, inp, out, ads, out=(proof, bl))
    assert not eng.debug_read_staging(0, n * 32).any()               # sk of all six pieces: wiped
    assert not eng.debug_read_staging(8, n * 32).any()               # and the blinding factors
    for lo in range(0, n, chunk):
        hi = lo + chunk
        cc, ss = eng.ietf_prove(vrfs.BANDERSNATCH, sk[lo:hi], inp[lo:hi], out[lo:hi], ads[lo:hi])
        assert np.array_equal(cc, c[lo:hi]) and np.array_equal(ss, s[lo:hi])
        pp, bb = eng.pedersen_prove(vrfs.BANDERSNATCH, sk[lo:hi], inp[lo:hi], out[lo:hi], ads[lo:hi])
        assert np.array_equal(pp, proof[lo:hi]) and np.array_equal(bb, bl[lo:hi])
    # the serialised proof form goes through the same pieces: it must be the encoding of the proofs above
    enc, bl2 = eng.pedersen_prove_compressed(vrfs.BANDERSNATCH, sk, inp, out, ads)
    assert np.array_equal(bl2, bl)
    enc_small, _ = eng.pedersen_prove_compressed(vrfs.BANDERSNATCH, sk[:chunk], inp[:chunk], out[:chunk], ads[:chunk])
    assert np.array_equal(enc_small, enc[:chunk])
    tail_small, _ = eng.pedersen_prove_compressed(vrfs.BANDERSNATCH, sk[-chunk:], inp[-chunk:], out[-chunk:], ads[-chunk:])
    assert np.array_equal(tail_small, enc[-chunk:])
    assert eng.pedersen_verify_compressed(vrfs.BANDERSNATCH, inp, out, enc, ads).all()
    # the first 1 024 items against the oracle
    co, so = O.ietf_prove(O.BANDERSNATCH, sk0, inp0, out0, ads[:base])
    assert np.array_equal(co, c[:base]) and np.array_equal(so, s[:base])
    # verifiers: corrupt a spread of items (in every piece), expect exactly those to fail
    bad = np.zeros(n, bool); bad[::997] = True; bad[-1] = True
    s2 = np.array(s); s2[bad, 0] ^= 1
    lib = _lib.load()
    assert lib.vrfs_host_register(C.c_void_p(s2.ctypes.data), C.c_size_t(s2.nbytes)) == _lib.OK
    try:
        ok = eng.ietf_verify(vrfs.BANDERSNATCH, pk, inp, out, c, s2, ads)
    finally:
        assert lib.vrfs_host_unregister(C.c_void_p(s2.ctypes.data)) == _lib.OK
    assert np.array_equal(ok.astype(bool), ~bad)
    proof2 = np.array(proof); proof2[bad, 200] ^= 1
    ok = eng.pedersen_verify(vrfs.BANDERSNATCH, inp, out, proof2, ads)
    assert np.array_equal(ok.astype(bool), ~bad)


@pytest.mark.gpu
def test_fixed_base_tables_are_built_on_first_use():
    import ark_ec_vrfs_b200 as vrfs
    with vrfs.Engine(0) as e:
        assert e.launch_count == 0                                   # nothing built at context creation
        sk, pk, inp, out = V.make_keys_inputs(O.ED25519, 8)
        e.output(vrfs.ED25519, sk, inp)                              # variable-base only: still no table
        a = e.launch_count
        e.ietf_prove(vrfs.ED25519, sk, inp, out)
        b = e.launch_count
        e.ietf_prove(vrfs.ED25519, sk, inp, out)
        assert (b - a) - (e.launch_count - b) == 2                   # the first prove built G's and B's table of this suite only


def test_wrapper_rejects_short_variable_length_lists():
    """a list of `ad` / data shorter than the batch must be a clean error, not an out-of-bounds read in the C side"""
    from ark_ec_vrfs_b200.engine import pack_var
    with pytest.raises(ValueError):
        pack_var([b"a", b"b"], 3)
    with pytest.raises(ValueError):
        pack_var((np.zeros(4, np.uint8), np.array([0, 2, 9], np.uint64)), 2)
    data, off = pack_var([b"ab", b"", b"cde"], 3)
    assert off.tolist() == [0, 2, 2, 5] and data.tobytes() == b"abcde"


# ---- several GPUs ---------------------------------------------------------------------------------------------------------
def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _srs(n, tag):
    import hashlib
    sc = np.frombuffer(b"".join(hashlib.sha256(tag + i.to_bytes(4, "little")).digest()[:8] for i in range(n)), np.uint8).reshape(n, 8)
    ks = np.zeros((n, 32), np.uint8); ks[:, :8] = sc
    return O.g1_mul_gen(ks)


@pytest.mark.gpu
@pytest.mark.parametrize("ndev", [1, 2, 4, 8])
def test_multi_device_context(ndev):
    """vrfs_ctx_create_multi: a verify batch sharded over the devices gives the single-device verdicts (ragged `ad`, a batch size
    that does not divide), the point-range MSM and the row-split ring commitment equal the single-device results"""
    if _gpus() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    import ark_ec_vrfs_b200 as vrfs
    n = 1000 + ndev
    w = V.make_ietf_proofs(O.BANDERSNATCH, n, "ragged")
    with vrfs.MultiEngine(list(range(ndev))) as m:
        ok, st = m.ietf_verify(vrfs.BANDERSNATCH, w["pk"], w["inp"], w["out"], w["c"], w["s"], w["ads"], status=True)
        assert np.array_equal(ok, w["expect"]) and np.array_equal(ok == 1, st == 0)
        N = 512
        bases = _srs(N, b"multi")
        rng = np.random.default_rng(ndev)
        sc = rng.integers(0, 256, size=(3 * N, 32), dtype=np.uint8); sc[:, 31] &= 0x3F
        h = m.msm_g1_prepare(bases)
        got = h.msm(sc, 3)
        assert np.array_equal(got, O.msm_g1(bases, sc, 3))
        assert np.array_equal(h.msm(sc, 3), got)                     # a second exchange (the other mailbox parity)
        assert np.array_equal(h.msm(sc[:N], 1), got[:1])
        with vrfs.Engine(0) as e:
            _, keys = e.secret_from_seed(vrfs.BANDERSNATCH, [b"mk-%d" % i for i in range(300)])
            tail, padding, keys = keys[260:], keys[259], keys[:259]
            part = N - 3 - len(tail) - 1
            h1 = e.msm_g1_prepare(bases)
            single = h1.ring_commit(keys, part, padding, tail, lagrange=True)
            h1.release()
        assert np.array_equal(h.ring_commit(keys, part, padding, tail), single)
        assert np.array_equal(h.ring_commit(keys[:5], part, padding, tail), vrfs_single_ring(bases, keys[:5], part, padding, tail))
        h.release()


def vrfs_single_ring(bases, keys, part, padding, tail):
    import ark_ec_vrfs_b200 as vrfs
    with vrfs.Engine(0) as e:
        h = e.msm_g1_prepare(bases)
        r = h.ring_commit(keys, part, padding, tail, lagrange=True)
        h.release()
    return r


@pytest.mark.gpu
def test_peer_group_of_one_rank(eng):
    """the device-side exchange with world = 1: the final kernel publishes to its own mailbox and folds"""
    N = 300
    bases = _srs(N, b"solo")
    rng = np.random.default_rng(3)
    sc = rng.integers(0, 256, size=(2 * N, 32), dtype=np.uint8); sc[:, 31] &= 0x3F
    eng.peer_connect(eng.peer_export(0, 1))
    assert eng.peer_world == 1
    h = eng.msm_g1_prepare(bases)
    exp = O.msm_g1(bases, sc, 2)
    for _ in range(3):
        assert np.array_equal(h.msm_allgather(sc, 2), exp)
    h.release()


def _peer_worker(rank, world, port, q):
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
    import torch
    import torch.distributed as dist
    import ark_ec_vrfs_b200 as vrfs
    from ark_ec_vrfs_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = vrfs.Engine(rank)
    connected = D.connect_peers(eng)
    N = 1 << 11
    bases = _srs(N, b"ipc")
    rng = np.random.default_rng(11)
    sc = rng.integers(0, 256, size=(3 * N, 32), dtype=np.uint8); sc[:, 31] &= 0x3F
    sh = D.ShardedPreparedBases(eng, bases)
    outs = [sh.msm(sc, 3) for _ in range(3)]
    _, keys = eng.secret_from_seed(vrfs.BANDERSNATCH, [b"pk-%d" % i for i in range(700)])
    tail, padding, keys = keys[600:], keys[599], keys[:599]
    part = N - 3 - len(tail) - 1
    rc = D.ShardedRingContext(eng, bases, part, padding, tail)
    ring = rc.verifier_key_commitment(keys)
    dist.barrier()
    q.put((rank, connected, sh.device_exchange, [o.tobytes() for o in outs], ring.tobytes()))
    rc.release(); sh.release(); eng.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_peer_exchange_across_processes(world):
    """one process per GPU: CUDA IPC mailboxes, the partials exchanged inside the MSM's last kernel; every rank gets the full
    commitment, equal to the oracle's MSM and to the single-GPU ring commitment"""
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    N = 1 << 11
    bases = _srs(N, b"ipc")
    rng = np.random.default_rng(11)
    sc = rng.integers(0, 256, size=(3 * N, 32), dtype=np.uint8); sc[:, 31] &= 0x3F
    exp = O.msm_g1(bases, sc, 3).tobytes()
    import ark_ec_vrfs_b200 as vrfs
    with vrfs.Engine(0) as e:
        _, keys = e.secret_from_seed(vrfs.BANDERSNATCH, [b"pk-%d" % i for i in range(700)])
    tail, padding, keys = keys[600:], keys[599], keys[:599]
    ring = vrfs_single_ring(bases, keys, N - 3 - len(tail) - 1, padding, tail).tobytes()
    for rank, connected, devx, outs, r in res:
        assert connected and devx, "the peer group did not come up on rank %d" % rank
        assert all(o == exp for o in outs) and r == ring
