#!/bin/bash
# Round-end GPU run: tests, bench, launch list, ncu full-set summaries (only the condensed CSVs are kept: gpurun_out/ is capped at 64 MiB).
# usage (on the GPU box, from the repo root): bash tools/run_round_profile.sh <tag>
tag=${1:-r1o}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; head -c 400 gpurun_out/${tag}_bench.json; echo
MSM_ONLY=1 timeout 300 python tools/bench_all.py --out gpurun_out/${tag}_bench_all_msm.json 2>&1 | grep "^msm" | cut -c1-120
timeout 200 python tools/msm_ring_columns.py > gpurun_out/${tag}_msm_ring_columns.log 2>&1
RING_BENCH_OUT=${tag}_ring_bench.json timeout 200 python tools/ring_bench.py > gpurun_out/${tag}_ring_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_b.log 2>&1
for logn in 17 11; do
  MSM_LOGN=$logn timeout 600 ncu --set full --clock-control none -k regex:k_msm -s 12 -c 12 -f -o /tmp/${tag}_msm_$logn python tools/msm_profile_run.py > gpurun_out/${tag}_msm_${logn}_ncu.log 2>&1
  ncu -i /tmp/${tag}_msm_$logn.ncu-rep --page raw --csv > gpurun_out/${tag}_msm_${logn}_raw.csv 2>/dev/null
done
ls -la gpurun_out | head -30
