#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/t_multi.log 2>&1; echo "multi rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench_n2 rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo "bench_n1 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_multi.log; tail -1 gpurun_out/bench_n2.log | cut -c1-400; tail -1 gpurun_out/bench_n1.log | cut -c1-200
