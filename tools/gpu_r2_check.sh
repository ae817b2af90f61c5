#!/bin/bash
# round-2 iteration flow: GPU suite (fault-injection tests in their own process, last), one bench line
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --deselect tests/test_gpu_faults.py ${PYTEST_ARGS} > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_faults.py -q -m gpu --timeout 600 > gpurun_out/t_faults.log 2>&1; echo "pytest_faults rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -12 gpurun_out/t_all.log | cut -c1-300; tail -12 gpurun_out/t_faults.log | cut -c1-300; tail -1 gpurun_out/bench.log | cut -c1-3000
