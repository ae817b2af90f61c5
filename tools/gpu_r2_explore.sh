#!/bin/bash
# round-2 exploratory captures of the round-1 engine: 64-channel conv at 512^2 (plain and with the in-kernel transform), launch lists at B=1 and B=8
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r2_c64_plain python tools/profile_conv.py --cin 64 --cout 64 --h 512 > gpurun_out/r2_ncu_c64_plain.log 2>&1
echo "c64 plain rc=$?"
CFB_TC_XFORM=2 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_tc_kernel<64, 2,' -c 2 -f -o gpurun_out/r2_c64_xf python tools/profile_forward.py --batch 4 > gpurun_out/r2_ncu_c64_xf.log 2>&1
echo "c64 xf rc=$?"
CFB_CUDA_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b1.csv python tools/profile_forward.py --batch 1 > gpurun_out/r2_pf_b1.log 2>&1
echo "launches b1 rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b8_r1engine.csv python tools/profile_forward.py --batch 8 > gpurun_out/r2_pf_b8.log 2>&1
echo "launches b8 rc=$?"
ls -la gpurun_out | tail -12
