#!/bin/bash
# Round-end GPU run: tests, bench, launch list, ncu full-set summaries (only the condensed CSVs are kept: gpurun_out/ is capped at 64 MiB).
# usage (on the GPU box, from the repo root): bash tools/run_round_profile.sh <tag>
tag=${1:-r2}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; head -c 300 gpurun_out/${tag}_bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; head -c 300 gpurun_out/${tag}_bench_reference.json; echo
MSM_ONLY=1 timeout 300 python tools/bench_all.py --out gpurun_out/${tag}_bench_all_msm.json 2>&1 | grep "^msm" | cut -c1-120
timeout 200 python tools/msm_ring_columns.py > gpurun_out/${tag}_msm_ring_columns.log 2>&1
timeout 200 python tools/pairing_bench.py > gpurun_out/${tag}_pairing_bench.log 2>&1
# launch list of the bench command (per-launch device times, cold cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --headline-only --no-cpu-baseline > gpurun_out/${tag}_ncu_b.log 2>&1
# full-set capture of the two lincomb kernels of one verify step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lincomb -s 8 -c 2 -f -o /tmp/${tag}_lincomb python bench.py --steps 1 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/${tag}_lincomb_ncu.log 2>&1
ncu -i /tmp/${tag}_lincomb.ncu-rep --page raw --csv > gpurun_out/${tag}_lincomb_raw.csv 2>/dev/null
# MSM: the second call of tools/msm_profile_run.py.  2^17 = bucket pipeline (3 preparation + 9 kernels per call), 2^11 = table mode (3 + 3 per call)
MSM_LOGN=17 timeout 600 ncu --set full --clock-control none -k regex:k_msm -s 12 -c 9 -f -o /tmp/${tag}_msm_17 python tools/msm_profile_run.py > gpurun_out/${tag}_msm_17_ncu.log 2>&1
ncu -i /tmp/${tag}_msm_17.ncu-rep --page raw --csv > gpurun_out/${tag}_msm_17_raw.csv 2>/dev/null
MSM_LOGN=11 timeout 600 ncu --set full --clock-control none -k regex:k_msm -s 6 -c 3 -f -o /tmp/${tag}_msm_11 python tools/msm_profile_run.py > gpurun_out/${tag}_msm_11_ncu.log 2>&1
ncu -i /tmp/${tag}_msm_11.ncu-rep --page raw --csv > gpurun_out/${tag}_msm_11_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:k_pairing_products_lanes -s 1 -c 1 -f -o /tmp/${tag}_pairing python tools/pairing_one.py > gpurun_out/${tag}_pairing_ncu.log 2>&1
ncu -i /tmp/${tag}_pairing.ncu-rep --page raw --csv > gpurun_out/${tag}_pairing_raw.csv 2>/dev/null
ls -la gpurun_out | tail -20
