#!/usr/bin/env python3
"""Secondary benchmarks (BASELINE.json configs 3-5): every suite's prove / verify / pedersen / h2c throughput and the
ring-commitment MSM, through the host-buffer C ABI (H2D/D2H included), one GPU.  Workloads are produced by the
engine itself at full size and cross-checked against the CPU oracle on a 2^10 subsample.
  python tools/bench_all.py [--logn 18] [--out gpurun_out/bench_all.json]"""
import argparse, hashlib, json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _scalars import fr_uniform
import ark_ec_vrfs_b200 as vrfs
import oracle_lib as O

ap = argparse.ArgumentParser(); ap.add_argument("--logn", type=int, default=18); ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_all.json"))
ap.add_argument("--msm-max-logn", type=int, default=17)
a = ap.parse_args()
n = 1 << a.logn
e = vrfs.Engine(0)
res = {"batch": n, "note": "host-buffer ABI calls (pageable numpy buffers), best of 3, includes H2D/D2H", "suites": {}}

DEV = {}          # kernel-only items/s of the most recent best() call (sum of the call's kernel times, CUDA events)


def best(f, reps=3):
    f(); ts = []
    e.enable_kernel_timing(True)
    for _ in range(reps):
        t = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t)
    DEV["kt"] = e.kernel_timings()          # (kernel, ms) of the timed calls, summed per kernel name
    DEV["ms"] = sum(ms for _, ms in DEV["kt"])
    e.enable_kernel_timing(False)
    return min(ts), r

sub = np.arange(0, n, n // 1024)
for suite, name in (() if os.environ.get("MSM_ONLY") else ((0, "bandersnatch"), (1, "ed25519"), (2, "secp256r1"))):
    seeds = [b"bench-sk" + i.to_bytes(8, "little") for i in range(n)]
    alphas = [i.to_bytes(8, "little") + bytes(24) for i in range(n)]
    r = {}; dev = {}
    t, (sk, pk) = best(lambda: e.secret_from_seed(suite, seeds)); r["secret_from_seed+public"] = n / t; dev["secret_from_seed+public"] = n / (DEV["ms"] * 1e-3);
    t, (inp, ok) = best(lambda: e.data_to_point(suite, alphas)); r["data_to_point"] = n / t; dev["data_to_point"] = n / (DEV["ms"] * 1e-3); assert ok.all()
    t, out = best(lambda: e.output(suite, sk, inp)); r["output"] = n / t; dev["output"] = n / (DEV["ms"] * 1e-3);
    t, (c, s) = best(lambda: e.ietf_prove(suite, sk, inp, out)); r["ietf_prove"] = n / t; dev["ietf_prove"] = n / (DEV["ms"] * 1e-3);
    t, okv = best(lambda: e.ietf_verify(suite, pk, inp, out, c, s)); r["ietf_verify"] = n / t; dev["ietf_verify"] = n / (DEV["ms"] * 1e-3); assert okv.all()
    t, (pr, bl) = best(lambda: e.pedersen_prove(suite, sk, inp, out)); r["pedersen_prove"] = n / t; dev["pedersen_prove"] = n / (DEV["ms"] * 1e-3);
    t, okp = best(lambda: e.pedersen_verify(suite, inp, out, pr)); r["pedersen_verify"] = n / t; dev["pedersen_verify"] = n / (DEV["ms"] * 1e-3); assert okp.all()
    pk_enc = e.point_encode(suite, pk); alphas_packed = vrfs.pack_var(alphas)
    t, (sig, sok) = best(lambda: e.ietf_sign_wire(suite, sk, alphas_packed)); r["ietf_sign_wire"] = n / t; dev["ietf_sign_wire"] = n / (DEV["ms"] * 1e-3); assert sok.all()
    t, (okw, beta) = best(lambda: e.ietf_verify_wire(suite, pk_enc, alphas_packed, sig)); r["ietf_verify_wire"] = n / t; dev["ietf_verify_wire"] = n / (DEV["ms"] * 1e-3); assert okw.all()
    ow, bw = O.ietf_verify_wire(suite, pk_enc[sub], [alphas[i] for i in sub], sig[sub]); assert ow.all() and np.array_equal(bw, beta[sub])
    # oracle cross-check on a subsample
    assert np.array_equal(O.data_to_point(suite, [alphas[i] for i in sub])[0], inp[sub])
    co, so = O.ietf_prove(suite, sk[sub], inp[sub], out[sub]); assert np.array_equal(co, c[sub]) and np.array_equal(so, s[sub])
    po, bo = O.pedersen_prove(suite, sk[sub], inp[sub], out[sub]); assert np.array_equal(po, pr[sub]) and np.array_equal(bo, bl[sub])
    res["suites"][name] = {k: round(v) for k, v in r.items()}
    res.setdefault("suites_kernel_only", {})[name] = {k: round(v) for k, v in dev.items()}
    print(name, "kernel-only", res["suites_kernel_only"][name], flush=True)
    print(name, res["suites"][name], flush=True)

# ring commitment MSM: bases k_i * G (62-bit multiples, cheap on the oracle), 3 random columns
R_BLS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
res["msm_g1_3col_ms"] = {}
rng = np.random.default_rng(5)
N = 1 << a.msm_max_logn
k = rng.integers(1, 2 ** 62, size=N, dtype=np.uint64); ks = np.zeros((N, 32), np.uint8); ks[:, :8] = k.view(np.uint8).reshape(N, 8)
bases_all = O.g1_mul_gen(ks)
for logn in range(10, a.msm_max_logn + 1):
    m = 1 << logn
    sc = fr_uniform(rng, 3 * m)
    t, outp = best(lambda: e.msm_g1(bases_all[:m], sc, 3))
    kt = DEV["kt"]
    if logn <= 12:
        assert np.array_equal(outp, O.msm_g1(bases_all[:m], sc, 3))
    h = e.msm_g1_prepare(bases_all[:m])
    tp, outp2 = best(lambda: h.msm(sc, 3))
    ktp = DEV["kt"]
    h.release()
    assert np.array_equal(outp, outp2)
    res["msm_g1_3col_ms"]["2^%d" % logn] = {"stateless_wall_ms": round(t * 1e3, 3), "stateless_kernels_ms": {a_: round(b_, 3) for a_, b_ in kt},
                                             "prepared_wall_ms": round(tp * 1e3, 3), "prepared_kernels_ms": {a_: round(b_, 3) for a_, b_ in ktp},
                                             "prepared_device_ms": round(sum(b_ for _, b_ in ktp), 3), "stateless_device_ms": round(sum(b_ for _, b_ in kt), 3)}
    print("msm 2^%d x3: stateless %.3f ms (device %.3f)  prepared %.3f ms (device %.3f)" % (logn, t * 1e3, sum(b_ for _, b_ in kt), tp * 1e3, sum(b_ for _, b_ in ktp)),
          {a_: round(b_, 2) for a_, b_ in ktp}, flush=True)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
