#!/bin/bash
# round-2 iteration flow: GPU suite (all failures listed, not -x), smoke, one bench line
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 ${PYTEST_ARGS} > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -25 gpurun_out/t_all.log; tail -1 gpurun_out/bench.log | cut -c1-1500
