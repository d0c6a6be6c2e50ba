#!/usr/bin/env python3
"""Point-range-sharded prepared MSM with the device-side exchange (the path bench.py's ring_kzg_msm_ms.sharded times), per size:
median wall ms of `msm_local` over the ranks (barrier before every call) and rank 0's kernel times.
   torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/msm_sharded_ab.py"""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _scalars import fr_uniform
import torch, torch.distributed as dist
import ark_ec_vrfs_b200 as vrfs
from ark_ec_vrfs_b200 import dist as D
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = vrfs.Engine(local)
assert D.connect_peers(eng)
eng.peer_set_timeout_ms(30000)
rng = np.random.default_rng(11)
gx = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
gy = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1
gen = np.zeros((1, 96), np.uint8); gen[0, :48] = np.frombuffer(gx.to_bytes(48, "little"), np.uint8); gen[0, 48:] = np.frombuffer(gy.to_bytes(48, "little"), np.uint8)
ks = np.zeros((2048, 32), np.uint8); ks[:, :8] = rng.integers(1, 2 ** 62, size=2048, dtype=np.uint64).view(np.uint8).reshape(2048, 8)
h1 = eng.msm_g1_prepare(gen)
small = np.concatenate([h1.msm(ks[i:i + 32], 32) for i in range(0, 2048, 32)]); h1.release()
res = {}
for logn in (11, 14, 17):
    n = 1 << logn
    bases = np.tile(small, (max(1, n // 2048), 1))[:n]
    sc = fr_uniform(rng, 3 * n)
    sh = D.ShardedPreparedBases(eng, bases)
    loc = sh.local_scalars(sc, 3)
    for _ in range(3): sh.msm_local(loc, 3)
    ts = []
    for _ in range(30):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        sh.msm_local(loc, 3)
        ts.append(time.perf_counter() - t0)
    eng.enable_kernel_timing(True); sh.msm_local(loc, 3); kt = dict(eng.kernel_timings()); eng.enable_kernel_timing(False)
    t = torch.tensor([sorted(ts)[len(ts) // 2]], dtype=torch.float64).cuda(); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["2^%d" % logn] = {"wall_ms_median": round(float(t[0]) * 1e3, 3), "kernels_rank0": {k: round(v, 3) for k, v in kt.items()}}
    sh.release()
if rank == 0: print(json.dumps(res))
eng.close(); dist.destroy_process_group()
