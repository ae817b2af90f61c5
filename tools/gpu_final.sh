#!/bin/bash
# round-end style flow: full GPU suite, smoke, both bench arms
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench_ref rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_all.log; tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/bench_ref.log | cut -c1-300; tail -1 gpurun_out/bench.log
