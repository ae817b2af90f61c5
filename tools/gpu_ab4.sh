#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/ab.log
for lib in codeformer_b200/ab/lib_*.so; do
  CFB_LIB=$PWD/$lib timeout 120 python tools/conv_ab.py 2>&1 | tail -1 >> gpurun_out/ab.log
done
CFB_PDL=0 timeout 120 python tools/conv_ab.py 2>&1 | tail -1 >> gpurun_out/ab.log
CFB_LIB=$PWD/codeformer_b200/ab/lib_prev.so timeout 120 python tools/conv_ab.py 2>&1 | tail -1 >> gpurun_out/ab.log
cat gpurun_out/ab.log
