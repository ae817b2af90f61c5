"""Summarise an ncu launch list (csv from --metrics gpu__time_duration.sum) into per-kernel time shares."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline='') as f:
    lines = [ln for ln in f if not ln.startswith('==')]
rd = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    name = re.sub(r'^void ', '', name).replace('cfb::', '')
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1)
    tot[name] += ns
    cnt[name] += 1
total = sum(tot.values())
print(f'| kernel | launches | total ms | share |\n|---|---:|---:|---:|')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f'| `{k}` | {cnt[k]} | {v / 1e6:.3f} | {100 * v / total:.1f}% |')
print(f'| **all** | {sum(cnt.values())} | {total / 1e6:.3f} | 100% |')
