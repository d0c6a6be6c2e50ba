#!/usr/bin/env python3
"""Run on the GPU box: integer-pipe microbenchmarks + a first timing of the verify path."""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ark_ec_vrfs_b200 as vrfs
import oracle_lib as O, vectors as V

e = vrfs.Engine(0)
res = {"nproc": os.cpu_count()}
names = {0: "IMAD.WIDE.U32", 1: "IMAD(lo)", 2: "montmul_chain1", 3: "montmul_chain2", 4: "IMAD.HI.U32", 5: "carry_chain_rows_reg", 6: "carry_chain_rows_imm", 9: "DFMA"}
for v in sorted(names):
    macs, mhz = e.measure_mac32_peak(v)
    res[names[v]] = {"Tmac_per_s": macs / 1e12, "sm_mhz_est": mhz, "mac_per_clk_per_sm": macs / (mhz * 1e6) / 148}
    print(names[v], res[names[v]], flush=True)
base = V.make_ietf_proofs(0, 4096, "empty")
for logn in (12, 16, 20):
    w = V.tile(base, (1 << logn) // 4096)
    for rep in range(3):
        t = time.time(); got = e.ietf_verify(0, w["pk"], w["inp"], w["out"], w["c"], w["s"], None); dt = time.time() - t
    assert np.array_equal(got, w["expect"])
    res["verify_2^%d" % logn] = {"s": dt, "per_s": (1 << logn) / dt}
    print("verify 2^%d: %.4fs  %.3f M/s (host buffers, pageable)" % (logn, dt, (1 << logn) / dt / 1e6), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
e.enable_kernel_timing(True)
got = e.ietf_verify(0, w["pk"], w["inp"], w["out"], w["c"], w["s"], None)
res["kernels_2^20_ms"] = e.kernel_timings(); print(res["kernels_2^20_ms"])
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe%s.json" % os.environ.get("PROBE_TAG", "")), "w"), indent=1)
