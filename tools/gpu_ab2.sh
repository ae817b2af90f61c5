#!/bin/bash
# A/B of library builds on one box (codeformer_b200/ab/lib_*.so): single-conv times, then bench lines for the builds named in $BENCH_LIBS
mkdir -p gpurun_out; : > gpurun_out/ab.log
for lib in codeformer_b200/ab/lib_*.so; do
  CFB_LIB=$PWD/$lib timeout 120 python tools/conv_ab.py 2>&1 | tail -1 >> gpurun_out/ab.log
done
CFB_PDL=0 timeout 120 python tools/conv_ab.py --small 2>&1 | tail -1 >> gpurun_out/ab.log
cat gpurun_out/ab.log
for tag in $BENCH_LIBS; do
  CFB_LIB=$PWD/codeformer_b200/ab/lib_$tag.so timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_$tag.log 2>&1
  echo "== $tag rc=$?"; tail -1 gpurun_out/bench_$tag.log | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'dominant ms', round(r['ms_per_launch'], 4), {k[:30]: round(v['ms_per_launch'], 4) for k, v in r['other_kernels'].items()}, d['clocks'])
"
done
CFB_CUDA_GRAPH=0 timeout 150 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b1.csv python tools/profile_forward.py --batch 1 > gpurun_out/pf_b1.log 2>&1; echo "launches b1 rc=$?"
python tools/summarize_launches.py gpurun_out/launches_b1.csv > gpurun_out/launch_summary_b1.md 2>&1; tail -1 gpurun_out/launch_summary_b1.md
timeout 200 bash tools/gpu_r2_vq.sh 2>&1 | tail -12
