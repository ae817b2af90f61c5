#!/bin/bash
# 2-GPU check of the shipped build (budget-tight): NCCL parity tests + second-device test, bench at N=1 and N=2 on the same box
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_faults.py::test_net_follows_its_device_and_second_device" -q -m gpu --timeout 250 > gpurun_out/t_multi.log 2>&1; echo "multi rc=$?"
tail -3 gpurun_out/t_multi.log | cut -c1-200
timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_n1.log 2>&1; echo "bench n1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"
for f in bench_n1 bench_n2; do tail -1 gpurun_out/$f.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('$f', round(d['value'], 1), round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d.get('multi_gpu'))
except Exception as e: print('$f parse failed', e)
"; done
