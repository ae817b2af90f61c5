#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_kernels.py -q -m gpu -s -k "config1 or batch_vs or variants or conv" > gpurun_out/t_e2e.log 2>&1; echo "tests rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_forward.py --batch 8 > gpurun_out/ncu_run.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_summary.md 2>&1
cat gpurun_out/summary.txt; grep -E "max-abs" gpurun_out/t_e2e.log | grep -v print; tail -3 gpurun_out/t_e2e.log; cut -c1-200 gpurun_out/bench.log | tail -1; head -9 gpurun_out/launch_summary.md
