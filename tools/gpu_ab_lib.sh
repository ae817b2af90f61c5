#!/bin/bash
# A/B of library builds on one box: codeformer_b200/ab/lib*.so, each with CFB_PDL=0 and 1 (older builds ignore the variable)
mkdir -p gpurun_out
summ() { tail -1 $1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
r = d['roofline']
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'u8', round(d['e2e_u8']['value'], 1), 'lat_b1', round(d['latency_b1_ms']['value'], 3), 'vq_ms', round(d['vq_micro']['ms'], 4), round(d['vq_micro'].get('ms_pipelined', 0), 4), 'vqae', round(d['vqae_b64']['faces_per_s'], 1), 'pn', round(d['parsenet_b8']['ms_per_step'],2), 'rrdb', round(d['rrdbnet_tile']['ms_per_tile'],2))
print('   dominant ms', round(r['ms_per_launch'], 4), {k[:40]: round(v['ms_per_launch'], 4) for k, v in r['other_kernels'].items()}, d['clocks'])
"; }
if [ -n "$RUN_TESTS" ]; then
  timeout 1200 python -m pytest tests -q -m gpu -x --timeout 900 --deselect tests/test_gpu_faults.py > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?"; tail -3 gpurun_out/t_all.log | cut -c1-300
fi
for lib in codeformer_b200/ab/lib*.so; do
  for pdl in 0 1; do
    case "$lib" in *head*) [ $pdl = 1 ] && continue;; esac
    tag=$(basename $lib .so)_pdl$pdl
    CFB_LIB=$PWD/$lib CFB_PDL=$pdl timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.log 2>&1
    echo "== $tag rc=$?"; summ gpurun_out/bench_$tag.log
  done
done
