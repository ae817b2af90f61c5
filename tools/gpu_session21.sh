#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -s -k "not block_by" > gpurun_out/t_e2e.log 2>&1; echo "tests rc=$?" >> gpurun_out/summary.txt
for l in 1 2 3 4; do
CFB_STREAM_LANES=$l timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l$l.log 2>&1; echo "bench lanes=$l rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; grep -E "max-abs" gpurun_out/t_e2e.log | grep -v print; tail -3 gpurun_out/t_e2e.log; for l in 1 2 3 4; do tail -1 gpurun_out/bench_l$l.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lanes', $l, d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['ms_per_launch'])"; done
