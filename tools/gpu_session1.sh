#!/bin/bash
# first GPU session: build check, kernel + e2e parity, smoke, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_e2e.py -x -q -m gpu -s > gpurun_out/t_e2e.log 2>&1; echo "e2e rc=$?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_kernels.log; tail -8 gpurun_out/t_e2e.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log
