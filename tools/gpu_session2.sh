#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python tools/tc_probe.py > gpurun_out/tc_probe.log 2>&1; echo "tc_probe rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -s -k "block_by_block and f32" > gpurun_out/t_blocks_f32.log 2>&1; echo "blocks_f32 rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -s -k "block_by_block and auto" > gpurun_out/t_blocks_auto.log 2>&1; echo "blocks_auto rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -s -k "not block_by_block" > gpurun_out/t_e2e.log 2>&1; echo "e2e rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/tc_probe.log | tail -25; tail -5 gpurun_out/t_kernels.log; grep -E "^\[|   (enc|gen|fuse|ft|quant)" gpurun_out/t_blocks_f32.log | head -80
