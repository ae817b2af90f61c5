#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 120 python tools/umma_probe.py > gpurun_out/umma_probe.log 2>&1; echo "umma_probe rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -s -k "config1 or batch_vs" > gpurun_out/t_e2e.log 2>&1; echo "e2e rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/umma_probe.log; grep max-abs gpurun_out/t_e2e.log | grep -v print; tail -3 gpurun_out/t_e2e.log
