#!/bin/bash
# A/B of library builds (single-conv times), then the GPU suite + bench on the default build
mkdir -p gpurun_out; : > gpurun_out/ab.log
for lib in codeformer_b200/ab/lib_*.so; do
  CFB_LIB=$PWD/$lib timeout 120 python tools/conv_ab.py 2>&1 | tail -1 >> gpurun_out/ab.log
done
CFB_PDL=0 timeout 120 python tools/conv_ab.py 2>&1 | tail -1 >> gpurun_out/ab.log
cat gpurun_out/ab.log
timeout 420 python -m pytest tests -q -m gpu --timeout 300 --deselect tests/test_gpu_faults.py > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?"
tail -3 gpurun_out/t_all.log | cut -c1-200
timeout 200 python -m pytest tests/test_gpu_faults.py -q -m gpu --timeout 180 > gpurun_out/t_faults.log 2>&1; echo "pytest_faults rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'u8', round(d['e2e_u8']['value'], 1), 'lat_b1', round(d['latency_b1_ms']['value'], 3), 'vq_ms', round(d['vq_micro']['ms'], 4), round(d['vq_micro'].get('ms_pipelined', 0), 4), 'vqae', round(d['vqae_b64']['faces_per_s'], 1), 'pn', round(d['parsenet_b8']['ms_per_step'],2), 'rrdb', round(d['rrdbnet_tile']['ms_per_tile'],2))
print('dominant ms', round(r['ms_per_launch'], 4), {k[:30]: round(v['ms_per_launch'], 4) for k, v in r['other_kernels'].items()}, d['clocks'])
"
