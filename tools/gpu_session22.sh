#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -s > gpurun_out/t_e2e.log 2>&1; echo "tests rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/bench_extra.py > gpurun_out/bench_extra.log 2>&1; echo "extra rc=$?" >> gpurun_out/summary.txt
CFB_CUDA_GRAPH=0 timeout 300 python tools/bench_extra.py > gpurun_out/bench_extra_nograph.log 2>&1; echo "extra_nograph rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "max-abs" gpurun_out/t_e2e.log | grep -v print; tail -3 gpurun_out/t_e2e.log; tail -1 gpurun_out/bench_extra.log; tail -1 gpurun_out/bench_extra_nograph.log
