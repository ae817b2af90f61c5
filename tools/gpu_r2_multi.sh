#!/bin/bash
# 2-GPU checks: NCCL parity tests, second-device test, scaling bench at N=1 and N=2 on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_faults.py::test_net_follows_its_device_and_second_device" -q -m gpu --timeout 600 > gpurun_out/t_multi.log 2>&1; echo "multi rc=$?"
tail -5 gpurun_out/t_multi.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_n1.log 2>&1; echo "bench n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --gather-chunks 1 > gpurun_out/bench_n2_c1.log 2>&1; echo "bench n2 chunks1 rc=$?"
for f in bench_n1 bench_n2 bench_n2_c1; do tail -1 gpurun_out/$f.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('$f', d['value'], d['ms_per_step'], d.get('multi_gpu'))
except Exception as e: print('$f parse failed', e)
"; done
