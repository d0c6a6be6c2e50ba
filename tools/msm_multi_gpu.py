#!/usr/bin/env python3
"""Ring-commitment MSM (3 columns, BLS12-381 G1) split by point range over the GPUs of one box:
   torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/msm_multi_gpu.py [--logn 17]
Every rank prepares its slice of the SRS once, then each commitment = prepared partial MSM + all-gather of 3 x 144 bytes
per rank (NCCL) + G-1 point additions.  Timed per call with a barrier on both sides, max over ranks."""
import argparse, json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _scalars import fr_uniform
import torch
import torch.distributed as dist
import ark_ec_vrfs_b200 as vrfs
from ark_ec_vrfs_b200 import dist as vd

ap = argparse.ArgumentParser(); ap.add_argument("--logn", type=int, nargs="+", default=[11, 14, 17]); ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
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
res = {}
for logn in a.logn:
    n = 1 << logn
    bases = np.tile(small, (max(1, n // 2048), 1))[:n]
    sc = fr_uniform(rng, 3 * n)
    sh = vd.ShardedPreparedBases(eng, bases, device=dev)
    out = sh.msm(sc, 3)
    if world > 1:                                     # every rank must hold the same commitment
        g = vd.gather_bytes(out, device=dev)
        assert all(np.array_equal(g[0], g[i]) for i in range(world))
    if rank == 0 and logn <= 14:                      # and it must equal the single-GPU result
        hfull = eng.msm_g1_prepare(bases); assert np.array_equal(out, hfull.msm(sc, 3)); hfull.release()
    ts = []
    for _ in range(a.reps):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        sh.msm(sc, 3)
        torch.cuda.synchronize(); dist.barrier(); ts.append(time.perf_counter() - t0)
    t = torch.tensor([sorted(ts)[len(ts) // 2]], dtype=torch.float64)
    if world > 1:
        t = t.cuda(); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["2^%d" % logn] = round(float(t[0]) * 1e3, 3)
    sh.release()
if rank == 0:
    print(json.dumps({"metric": "ring_kzg_commitment_msm_ms", "n_gpus": world, "columns": 3, "ms_median_of_%d" % a.reps: res,
                      "what": "prepared SRS slice per rank, partial MSM + all-gather(3 x 144 B per rank) + fold; wall clock incl. host buffers, barrier on both sides, max over ranks"}))
eng.close()
dist.destroy_process_group()
