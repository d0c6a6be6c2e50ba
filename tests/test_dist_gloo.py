"""world_size-2 gloo test of the multi-GPU host logic on CPU: index-range sharding covers the
batch exactly once, per-rank verdicts concatenate to the single-rank result, and the 144-byte MSM partials
survive the all-gather.  The per-rank compute is the CPU oracle here (the CUDA path needs a GPU); what is
under test is ark_ec_vrfs_b200/dist.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle_lib as O
    import vectors as V
    from ark_ec_vrfs_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 37                                                    # deliberately not divisible by the world size
    w = V.make_ietf_proofs(O.BANDERSNATCH, n, "ragged")
    lo, hi = D.shard_range(n, rank, world)
    pk, inp, out, c, s = D.shard_arrays([w["pk"], w["inp"], w["out"], w["c"], w["s"]], rank, world)
    ads = D.shard_var(w["ads"], rank, world)
    local = O.ietf_verify(O.BANDERSNATCH, pk, inp, out, c, s, ads, nthreads=1)
    padded = np.zeros(n, np.uint8); padded[lo:hi] = local     # ranks write disjoint ranges
    allv = D.gather_bytes(padded)
    merged = allv.sum(axis=0).astype(np.uint8)
    part = np.full((3, 144), rank + 1, np.uint8)
    parts = D.gather_bytes(part)
    dist.barrier()
    if rank == 0:
        q.put((merged.tolist(), w["expect"].tolist(), parts.shape, [int(parts[r, 0, 0]) for r in range(world)]))
    dist.destroy_process_group()


def test_sharded_verify_and_partial_gather_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    merged, expect, shape, tags = q.get(timeout=180)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert merged == expect and 0 < sum(expect) < len(expect)
    assert tuple(shape) == (2, 3, 144) and tags == [1, 2]


@pytest.mark.parametrize("n,world", [(0, 1), (1, 8), (37, 2), (1 << 20, 8), (1000003, 4)])
def test_shard_ranges_partition(n, world):
    from ark_ec_vrfs_b200 import dist as D
    r = [D.shard_range(n, g, world) for g in range(world)]
    assert r[0][0] == 0 and r[-1][1] == n
    assert all(r[g][1] == r[g + 1][0] for g in range(world - 1))
    sizes = [hi - lo for lo, hi in r]
    assert max(sizes) - min(sizes) <= 1


def test_ring_column_rows_partition_the_columns():
    """dist.ring_column_rows: the row slices of all ranks concatenate to the full fixed columns (keys, padding, tail, zero rows;
    selector on the key slots) for ragged splits"""
    from ark_ec_vrfs_b200 import dist as D
    rng = np.random.default_rng(2)
    n, part = 64, 40
    keys = rng.integers(0, 256, size=(7, 64), dtype=np.uint8); tail = rng.integers(0, 256, size=(10, 64), dtype=np.uint8)
    padding = rng.integers(0, 256, size=64, dtype=np.uint8)
    full = D.ring_column_rows(0, n, part, keys, padding, tail)
    pts = np.concatenate([full[0], full[1]], axis=1)
    assert np.array_equal(pts[:7], keys) and np.array_equal(pts[7:part], np.tile(padding, (part - 7, 1)))
    assert np.array_equal(pts[part:part + 10], tail) and not pts[part + 10:].any()
    assert full[2][:, 0].tolist() == [1] * part + [0] * (n - part) and not full[2][:, 1:].any()
    for world in (1, 2, 3, 8):
        parts = [D.ring_column_rows(*D.shard_range(n, g, world), part, keys, padding, tail) for g in range(world)]
        assert np.array_equal(np.concatenate(parts, axis=1), full)
