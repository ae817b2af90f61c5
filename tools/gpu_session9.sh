#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
M="sm__pipe_tensor_subpipe_hmma_cycles_active,sm__inst_executed_pipe_tensor,sm__cycles_elapsed.max,sm__cycles_active.avg,gpu__time_duration.sum"
for cfg in "128 128 256" "64 64 512"; do
 set -- $cfg
 for halo in 1 0; do
  CFB_TC_HALO=$halo timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/prof_c$1_halo$halo python tools/profile_conv.py --cin $1 --cout $2 --h $3 > gpurun_out/ncu_c$1_halo$halo.log 2>&1; echo "ncu c$1 halo$halo rc=$?" >> gpurun_out/summary.txt
  ncu -i gpurun_out/prof_c$1_halo$halo.ncu-rep --page raw --csv > gpurun_out/prof_c$1_halo$halo.raw.csv 2>/dev/null
 done
done
cat gpurun_out/summary.txt; ls -la gpurun_out/*.ncu-rep
