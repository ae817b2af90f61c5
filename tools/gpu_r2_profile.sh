#!/bin/bash
# round-2 captures of the shipped engine: launch list of one forward at B=8, ncu --set full of the transform convs
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b8.csv python tools/profile_forward.py --batch 8 > gpurun_out/r2_pf_b8.log 2>&1
echo "launches b8 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r2_c64_xf python tools/profile_conv.py --cin 64 --cout 64 --h 512 --xf 1 > gpurun_out/r2_ncu_c64_xf.log 2>&1
echo "c64 xf rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r2_c128_xf python tools/profile_conv.py --cin 128 --cout 128 --h 256 --xf 1 > gpurun_out/r2_ncu_c128_xf.log 2>&1
echo "c128 xf rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_b8.csv > gpurun_out/r2_launch_summary_b8.md
head -40 gpurun_out/r2_launch_summary_b8.md
