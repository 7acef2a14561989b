"""Print the attention launch durations (us) and the block total of ncu launch lists: tools/attn_times.py p27 p28"""
import csv, sys
for tag in sys.argv[1:]:
    rows = list(csv.reader(open(f'gpurun_out/launches_block_{tag}.csv')))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    h = rows[hi]; k = h.index('Kernel Name'); v = h.index('Metric Value')
    at = [float(r[v]) / 1000 for r in rows[hi + 1:] if len(r) > v and 'attention' in r[k]]
    tot = sum(float(r[v]) for r in rows[hi + 1:] if len(r) > v) / 1000
    print(tag, [round(x, 1) for x in at], round(tot, 1))
