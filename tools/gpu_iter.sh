#!/bin/bash
# iteration flow: stamps, GPU suite without the fault tests, bench (no CPU baseline)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python tools/tc_stamps.py > gpurun_out/stamps.log 2>&1; echo "stamps rc=$?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu -x --timeout 900 --deselect tests/test_gpu_faults.py ${PYTEST_ARGS} > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -4 gpurun_out/t_all.log | cut -c1-300; tail -1 gpurun_out/bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
r = d['roofline']
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'u8', round(d['e2e_u8']['value'], 1), 'lat_b1', round(d['latency_b1_ms']['value'], 3), 'vq_ms', round(d['vq_micro']['ms'], 4), round(d['vq_micro'].get('ms_pipelined', 0), 4), 'vqae', round(d['vqae_b64']['faces_per_s'], 1))
print('dominant ms', round(r['ms_per_launch'], 4), 'frac', round(r['frac'], 4), {k[:40]: round(v['ms_per_launch'], 4) for k, v in r['other_kernels'].items()})
print('parsenet', d.get('parsenet_b8', {}).get('ms_per_step'), 'rrdb', d.get('rrdbnet_tile', {}).get('ms_per_tile'), 'clocks', d['clocks'])
"
