#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -s -k "vqauto or config1" > gpurun_out/t_e2e.log 2>&1; echo "e2e rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/bench_extra.py > gpurun_out/bench_extra.log 2>&1; echo "bench_extra rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_forward.py --batch 8 > gpurun_out/ncu_run.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_summary.md 2>&1
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_kernels.log; tail -4 gpurun_out/t_e2e.log; cat gpurun_out/bench_extra.log; head -12 gpurun_out/launch_summary.md
