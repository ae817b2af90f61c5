#!/bin/bash
# A/B on one box: env var $1 with values 0 / 1 -> stamps + bench for each
mkdir -p gpurun_out
for v in 0 1; do
  env $1=$v timeout 300 python tools/tc_stamps.py > gpurun_out/stamps_$v.log 2>&1
  env $1=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.log 2>&1
  echo "== $1=$v"; grep -E "^---" gpurun_out/stamps_$v.log | cut -c1-90
  tail -1 gpurun_out/bench_$v.log | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
r = d['roofline']
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'u8', round(d['e2e_u8']['value'], 1), 'lat_b1', round(d['latency_b1_ms']['value'], 3), 'vq_ms', round(d['vq_micro']['ms'], 4), round(d['vq_micro'].get('ms_pipelined', 0), 4), 'vqae', round(d['vqae_b64']['faces_per_s'], 1))
print('dominant ms', round(r['ms_per_launch'], 4), {k[:40]: round(v['ms_per_launch'], 4) for k, v in r['other_kernels'].items()}, d['clocks'])
"
done
