#!/bin/bash
# round-2 closing flow: A/B of the two candidate builds (single-conv times) -> the faster one runs the GPU suite, smoke, bench
# and the profile captures (launch list at B=8; ncu --set full of the two transform convs and of the VQ kernel)
mkdir -p gpurun_out; : > gpurun_out/ab.log; rm -f gpurun_out/summary.txt
for t in E G; do CFB_LIB=$PWD/codeformer_b200/ab/lib_$t.so timeout 120 python tools/conv_ab.py 2>&1 | tail -1 >> gpurun_out/ab.log; done
cat gpurun_out/ab.log
WIN=$(python - <<'PY'
import re
best = None
for ln in open('gpurun_out/ab.log'):
    m = re.match(r'\[lib_(\w+)\.so', ln)
    v = [float(x) for x in re.findall(r': ([0-9.]+)', ln)]
    if m and len(v) >= 3:
        s = v[0] * 20 + v[2] * 8 + v[1] * 4          # weights ~ launches per forward of the three big families
        if best is None or s < best[0]:
            best = (s, m.group(1))
print(best[1] if best else 'E')
PY
)
echo "winner $WIN" | tee gpurun_out/winner.txt
LOSE=$([ "$WIN" = E ] && echo G || echo E)
export CFB_LIB=$PWD/codeformer_b200/ab/lib_$WIN.so
timeout 420 python -m pytest tests -q -m gpu --timeout 300 --deselect tests/test_gpu_faults.py > gpurun_out/t_all.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 200 python -m pytest tests/test_gpu_faults.py -q -m gpu --timeout 180 > gpurun_out/t_faults.log 2>&1; echo "pytest_faults rc=$?" >> gpurun_out/summary.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
CFB_LIB=$PWD/codeformer_b200/ab/lib_$LOSE.so timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_other.log 2>&1; echo "bench_other($LOSE) rc=$?" >> gpurun_out/summary.txt
timeout 150 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b8.csv python tools/profile_forward.py --batch 8 > gpurun_out/pf_b8.log 2>&1; echo "launches b8 rc=$?" >> gpurun_out/summary.txt
python tools/summarize_launches.py gpurun_out/launches_b8.csv > gpurun_out/launch_summary_b8.md 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r2f_c128_xf python tools/profile_conv.py --cin 128 --cout 128 --h 256 --xf 1 > gpurun_out/ncu_c128.log 2>&1; echo "ncu c128 rc=$?" >> gpurun_out/summary.txt
timeout 150 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r2f_c64_xf python tools/profile_conv.py --cin 64 --cout 64 --h 512 --xf 1 > gpurun_out/ncu_c64.log 2>&1; echo "ncu c64 rc=$?" >> gpurun_out/summary.txt
timeout 120 ncu --set full --clock-control none --import-source on -k regex:vq_fused -s 2 -c 1 -f -o gpurun_out/r2f_vq_fused python tools/profile_vq.py > gpurun_out/ncu_vq.log 2>&1; echo "ncu vq rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_all.log | cut -c1-200; tail -2 gpurun_out/t_faults.log | cut -c1-200; tail -1 gpurun_out/smoke.log
for f in bench bench_other; do tail -1 gpurun_out/$f.log | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'u8', round(d['e2e_u8']['value'], 1), 'dominant ms', round(r['ms_per_launch'], 4), {k[:30]: round(v['ms_per_launch'], 4) for k, v in r['other_kernels'].items()}, d['clocks'])
if 'latency_b1_ms' in d: print('lat_b1', round(d['latency_b1_ms']['value'], 3), 'vq_ms', round(d['vq_micro']['ms'], 4), round(d['vq_micro'].get('ms_pipelined', 0), 4), 'vqae', round(d['vqae_b64']['faces_per_s'], 1), 'pn', round(d['parsenet_b8']['ms_per_step'],2), 'rrdb', round(d['rrdbnet_tile']['ms_per_tile'],2))
"; done
