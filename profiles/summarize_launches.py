"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:70]
    v = float(row['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(row['Metric Unit'], 1e-6)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f'# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot:.3f} ms (ncu per-launch times: cold cache, serialised)')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{v[1]:10.3f} ms {100 * v[1] / tot:5.1f}% {v[0]:6d}  {k}')
