#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for c in 4 8 16; do
CFB_TC_CHUNK=$c timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -s -k "config1 or batch_vs" > gpurun_out/t_c$c.log 2>&1; echo "tests chunk=$c rc=$?" >> gpurun_out/summary.txt
CFB_TC_CHUNK=$c timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c$c.log 2>&1; echo "bench chunk=$c rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
for c in 4 8 16; do grep -E "max-abs" gpurun_out/t_c$c.log | grep -v print | head -2; tail -1 gpurun_out/bench_c$c.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk', $c, round(d['value'],1), round(d['roofline']['ms_per_launch'],4))"; done
