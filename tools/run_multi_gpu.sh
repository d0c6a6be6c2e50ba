for g in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29600+g)) tools/msm_multi_gpu.py 2>/dev/null | grep '^{' > gpurun_out/r1o_msm_multi_gpu_$g.json; cat gpurun_out/r1o_msm_multi_gpu_$g.json | cut -c1-200
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r1o_bench_8gpu.json; head -c 500 gpurun_out/r1o_bench_8gpu.json
