#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -s -k "config1 or batch_vs or variants" > gpurun_out/t_e2e.log 2>&1; echo "e2e rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/profile_forward.py --batch 1 > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/prof_final_c128 python tools/profile_conv.py --cin 128 --cout 128 --h 256 > gpurun_out/ncu_full.log 2>&1; echo "ncu_full rc=$?" >> gpurun_out/summary.txt
ncu -i gpurun_out/prof_final_c128.ncu-rep --page raw --csv > gpurun_out/prof_final_c128.raw.csv 2>/dev/null
cat gpurun_out/summary.txt; grep -E "max-abs" gpurun_out/t_e2e.log | grep -v print; tail -3 gpurun_out/t_e2e.log; cut -c1-200 gpurun_out/bench.log | tail -1; tail -6 gpurun_out/memcheck.log
