#!/bin/bash
# state check (budget-tight): GPU suite, faults in their own process, smoke, bench, launch list at B=8
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 420 python -m pytest tests -q -m gpu --timeout 300 --deselect tests/test_gpu_faults.py --durations=10 > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 200 python -m pytest tests/test_gpu_faults.py -q -m gpu --timeout 180 > gpurun_out/t_faults.log 2>&1; echo "pytest_faults rc=$?" >> gpurun_out/summary.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 150 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b8.csv python tools/profile_forward.py --batch 8 > gpurun_out/pf_b8.log 2>&1; echo "launches b8 rc=$?" >> gpurun_out/summary.txt
python tools/summarize_launches.py gpurun_out/launches_b8.csv > gpurun_out/launch_summary_b8.md 2>&1
cat gpurun_out/summary.txt; tail -22 gpurun_out/t_all.log | cut -c1-200; tail -5 gpurun_out/t_faults.log | cut -c1-200; tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/bench.log | cut -c1-5000
