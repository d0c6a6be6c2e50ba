#!/usr/bin/env python3
"""Ring commitment split by rows over the GPUs of one box (SURVEY 8e + 8f-2, Lagrange-basis SRS):
   torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/ring_multi_gpu.py
dist.ShardedRingContext: every rank prepares its rows of the SRS once; a commitment = host-built column rows + prepared partial MSM +
all-gather of 3 x 144 bytes per rank + fold.  Checked against the single-GPU vrfs_ring_commit on rank 0; wall clock, max over ranks."""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import ark_ec_vrfs_b200 as vrfs
from ark_ec_vrfs_b200 import dist as vd

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29534")
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    dist.init_process_group("gloo", rank=0, world_size=1)
dev = torch.device("cuda", local) if world > 1 else None
eng = vrfs.Engine(local)
rng = np.random.default_rng(11)                      # same seed on every rank: identical inputs
gx = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
gy = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1
gen = np.zeros((1, 96), np.uint8); gen[0, :48] = np.frombuffer(gx.to_bytes(48, "little"), np.uint8); gen[0, 48:] = np.frombuffer(gy.to_bytes(48, "little"), np.uint8)
ks = np.zeros((2048, 32), np.uint8); ks[:, :8] = rng.integers(1, 2 ** 62, size=2048, dtype=np.uint64).view(np.uint8).reshape(2048, 8)
h1 = eng.msm_g1_prepare(gen)
small = np.concatenate([h1.msm(ks[i:i + 32], 32) for i in range(0, 2048, 32)]); h1.release()
_, pk_all = eng.secret_from_seed(vrfs.BANDERSNATCH, [b"ring-key-%d" % i for i in range((1 << 16) + 300)])
res = {}
for logn in (11, 14, 17):
    n = 1 << logn
    srs = np.tile(small, (max(1, n // 2048), 1))[:n]
    keys, padding, tail = pk_all[: n // 2], pk_all[1 << 16], pk_all[(1 << 16) + 1:(1 << 16) + 254]
    part = n - 3 - len(tail) - 1
    ctx = vd.ShardedRingContext(eng, srs, part, padding, tail, device=dev)
    out = ctx.verifier_key_commitment(keys)
    if rank == 0:
        h = eng.msm_g1_prepare(srs); assert np.array_equal(out, h.ring_commit(keys, part, padding, tail, lagrange=True)); h.release()
    ts = []
    for _ in range(10):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        ctx.verifier_key_commitment(keys)
        torch.cuda.synchronize(); dist.barrier(); ts.append(time.perf_counter() - t0)
    t = torch.tensor([sorted(ts)[len(ts) // 2]], dtype=torch.float64)
    if world > 1:
        t = t.cuda(); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["2^%d" % logn] = round(float(t[0]) * 1e3, 3)
    ctx.release()
if rank == 0:
    print(json.dumps({"metric": "ring_commitment_ms", "n_gpus": world, "ms_median_of_10": res,
                      "what": "dist.ShardedRingContext: ring of N/2 keys, padded, 253-row tail, Lagrange SRS rows per rank; equal to the single-GPU vrfs_ring_commit (asserted)"}))
eng.close()
dist.destroy_process_group()
