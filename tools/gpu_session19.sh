#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for cfg in "128 128 256" "64 64 512"; do
 set -- $cfg
 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/prof_res_c$1 python tools/profile_conv.py --cin $1 --cout $2 --h $3 --residual 1 > gpurun_out/ncu_res_c$1.log 2>&1; echo "ncu res c$1 rc=$?" >> gpurun_out/summary.txt
 ncu -i gpurun_out/prof_res_c$1.ncu-rep --page source --csv > gpurun_out/prof_res_c$1.src.csv 2>/dev/null
 ncu -i gpurun_out/prof_res_c$1.ncu-rep --page raw --csv > gpurun_out/prof_res_c$1.raw.csv 2>/dev/null
done
rm -f gpurun_out/prof_res_*.ncu-rep
cat gpurun_out/summary.txt
