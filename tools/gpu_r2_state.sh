#!/bin/bash
# state check: GPU suite (faults last, own process), one bench line with every extra key, launch lists at B=1 and B=8
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 --deselect tests/test_gpu_faults.py --durations=15 > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_faults.py -q -m gpu --timeout 600 > gpurun_out/t_faults.log 2>&1; echo "pytest_faults rc=$?" >> gpurun_out/summary.txt
CFB_CUDA_GRAPH=0 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b1.csv python tools/profile_forward.py --batch 1 > gpurun_out/pf_b1.log 2>&1; echo "launches b1 rc=$?" >> gpurun_out/summary.txt
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b8.csv python tools/profile_forward.py --batch 8 > gpurun_out/pf_b8.log 2>&1; echo "launches b8 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -25 gpurun_out/t_all.log | cut -c1-200; tail -5 gpurun_out/t_faults.log | cut -c1-200; tail -1 gpurun_out/bench.log | cut -c1-4000
