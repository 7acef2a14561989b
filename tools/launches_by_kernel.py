"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel:
   python tools/launches_by_kernel.py gpurun_out/launches_bench_TAG.csv profiles/TAG_launches_bench_C2_by_kernel.csv"""
import csv, re, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]; k = h.index('Kernel Name'); v = h.index('Metric Value'); u = h.index('Metric Unit')
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= v: continue
    name = re.sub(r'^void ', '', r[k]); name = name.split('(')[0]
    t = float(r[v].replace(',', ''))
    t = t / 1000.0 if r[u].startswith('ns') else t
    agg[name][0] += 1; agg[name][1] += t
tot = sum(x[1] for x in agg.values())
with open(sys.argv[2], 'w', newline='') as f:
    w = csv.writer(f); w.writerow(['kernel', 'launches', 'total_us', 'share'])
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        w.writerow([n, c, round(t, 1), round(t / tot, 4)])
print('total_us', round(tot, 1))
