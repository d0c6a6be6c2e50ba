"""Multi-GPU plumbing: one process per GPU under torch.distributed.

VRF batches are independent items: rank g takes the contiguous index range [g*n/G, (g+1)*n/G) and there is
NO data-path collective.  The MSM splits by point range; each rank produces one projective partial per
column (144 bytes).  On GPU boxes the partials are exchanged INSIDE the MSM's last kernel: `connect_peers`
all-gathers the ranks' 64-byte CUDA IPC mailbox handles once (host side, torch.distributed), after which
`ShardedPreparedBases.msm` / `ShardedRingContext.verifier_key_commitment` are single C-ABI calls whose final
kernel stores the partial into every peer's device memory over NVLink, waits for the peers' flags and folds
(vrfs_msm_g1_prepared_allgather / vrfs_ring_commit_rows_allgather).  Without a peer group (the gloo CPU
tests, or GPUs without peer access) the partials go through a host all-gather and `Engine.g1_sum_partials`."""
import numpy as np


def shard_range(n, rank, world):
    """contiguous, balanced, covering: the ranges of ranks 0..world-1 partition [0, n)"""
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_arrays(arrays, rank, world):
    n = len(arrays[0])
    lo, hi = shard_range(n, rank, world)
    return [a[lo:hi] for a in arrays]


def shard_var(items, rank, world):
    if items is None:
        return None
    lo, hi = shard_range(len(items), rank, world)
    return items[lo:hi]


def gather_bytes(local: np.ndarray, group=None, device=None):
    """all-gather equal-sized uint8 arrays; returns (world, ...) on every rank"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return np.stack([o.cpu().numpy() for o in outs])


def connect_peers(engine, group=None):
    """one-time setup of the device-side exchange: every rank exports its mailbox, the handles are all-gathered over
    torch.distributed (host objects), every rank maps its peers.  Returns True when the peer group is up; False (and the
    callers below fall back to the host all-gather) when the GPUs cannot map each other."""
    import torch.distributed as dist
    from ._lib import VrfsError
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    h = engine.peer_export(rank, world)
    handles = [None] * world
    dist.all_gather_object(handles, h.tobytes(), group=group)
    ok = True
    try:
        engine.peer_connect(np.frombuffer(b"".join(handles), np.uint8))
    except VrfsError:
        ok = False
    flags = [None] * world
    dist.all_gather_object(flags, ok, group=group)          # all or nothing: a collective needs every rank
    return all(flags)


def msm_g1_sharded(engine, bases, scalars, n_columns, group=None, device=None):
    """each rank passes the FULL inputs; computes its point range; all ranks return the same affine result"""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = len(bases)
    lo, hi = shard_range(n, rank, world)
    sc = np.asarray(scalars, np.uint8).reshape(n_columns, n, 32)[:, lo:hi].reshape(-1, 32)
    part = engine.msm_g1_partial(bases[lo:hi], sc, n_columns)
    parts = gather_bytes(part, group, device)
    return engine.g1_sum_partials(parts, n_columns)


class ShardedPreparedBases:
    """RingContext analogue on G GPUs: rank g prepares the SRS slice [g*n/G, (g+1)*n/G) once; every commitment is then one
    prepared partial MSM per rank, an all-gather of 144 bytes per column per rank and G-1 point additions."""

    def __init__(self, engine, bases, group=None, device=None):
        import torch.distributed as dist
        self.engine, self.group, self.device = engine, group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n = len(bases)
        self.lo, self.hi = shard_range(self.n, self.rank, self.world)
        self.handle = engine.msm_g1_prepare(bases[self.lo:self.hi])
        self.device_exchange = engine.peer_world == self.world      # connect_peers was called for this group

    def local_scalars(self, scalars, n_columns=1):
        """this rank's slice of column-major scalars (what a caller that already shards its data would hold)"""
        return np.ascontiguousarray(np.asarray(scalars, np.uint8).reshape(n_columns, self.n, 32)[:, self.lo:self.hi]).reshape(-1, 32)

    def msm_local(self, local_scalars, n_columns=1):
        """COLLECTIVE: the full commitments from each rank's own slice of the scalars"""
        if self.device_exchange:
            return self.handle.msm_allgather(local_scalars, n_columns)
        part = self.handle.msm_partial(local_scalars, n_columns)
        return self.engine.g1_sum_partials(gather_bytes(part, self.group, self.device), n_columns)

    def msm(self, scalars, n_columns=1):
        sc = self.local_scalars(scalars, n_columns)
        if self.device_exchange:
            return self.handle.msm_allgather(sc, n_columns)
        part = self.handle.msm_partial(sc, n_columns)
        parts = gather_bytes(part, self.group, self.device)
        return self.engine.g1_sum_partials(parts, n_columns)

    def release(self):
        self.handle.release()


def ring_column_rows(lo, hi, keyset_part_size, keys, padding, tail):
    """rows [lo, hi) of the ring's fixed columns (xs | ys | selector, the layout of vrfs_ring_fixed_columns) as (3, hi - lo, 32)
    canonical LE values - the host-side statement of what k_ring_columns writes for a rank's slice (used by the tests)"""
    keys = np.asarray(keys, np.uint8).reshape(-1, 64); tail = np.asarray(tail, np.uint8).reshape(-1, 64)
    padding = np.asarray(padding, np.uint8).reshape(64)
    rows = np.zeros((hi - lo, 64), np.uint8)
    idx = np.arange(lo, hi)
    k = idx < len(keys)
    rows[k] = keys[idx[k]]
    rows[(~k) & (idx < keyset_part_size)] = padding
    t = (idx >= keyset_part_size) & (idx < keyset_part_size + len(tail))
    rows[t] = tail[idx[t] - keyset_part_size]
    sel = np.zeros((hi - lo, 32), np.uint8); sel[idx < keyset_part_size, 0] = 1
    return np.stack([rows[:, :32], rows[:, 32:], sel])


class ShardedRingContext:
    """`RingContext` up to the verifier key's commitment on G GPUs with a Lagrange-basis SRS: rank g prepares the bases of the rows
    [g*N/G, (g+1)*N/G), builds those rows of the fixed columns on its GPU (vrfs_ring_commit_rows_partial), and the commitment is one partial MSM per rank, an all-gather of
    3 x 144 bytes per rank and G-1 point additions (every rank returns the same three points)."""

    def __init__(self, engine, srs_lagrange, keyset_part_size, padding, tail, group=None, device=None):
        self.bases = ShardedPreparedBases(engine, srs_lagrange, group, device)
        self.keyset_part_size, self.padding, self.tail = int(keyset_part_size), padding, tail

    def verifier_key_commitment(self, public_keys):
        b = self.bases
        keys = np.asarray(public_keys, np.uint8).reshape(-1, 64)
        if b.device_exchange:
            return b.handle.ring_commit_rows_allgather(b.lo, self.keyset_part_size, len(keys), keys[b.lo:b.hi], self.padding, self.tail)
        part = b.handle.ring_commit_rows_partial(b.lo, self.keyset_part_size, len(keys), keys[b.lo:b.hi], self.padding, self.tail)
        return b.engine.g1_sum_partials(gather_bytes(part, b.group, b.device), 3)

    def release(self):
        self.bases.release()
