"""Summarise the SASS source page of an ncu report exported with
   ncu -i rep --page source --csv --print-source sass > src.csv
usage: tools/ncu_stalls.py src.csv [top]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S = ix['# Samples']; E = ix['Instructions Executed']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[S] or 0) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = defaultdict(int)
for r in data:
    for h in stalls: agg[h] += int(r[ix[h]] or 0)
print('overall:', [(h, v) for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]])
by = defaultdict(lambda: [0, 0, defaultdict(int)])
for r in data:
    e = int(r[E] or 0); sm = int(r[S] or 0)
    by[e][0] += sm; by[e][1] += 1
    for h in stalls: by[e][2][h] += int(r[ix[h]] or 0)
print('by execution count (code region):')
for e, (sm, n, st) in sorted(by.items(), key=lambda x: -x[1][0])[:10]:
    print(' ', e, sm, n, sorted(st.items(), key=lambda x: -x[1])[:4])
print('top instructions:')
for r in sorted(data, key=lambda r: -int(r[S] or 0))[:top_n]:
    st = {h: int(r[ix[h]] or 0) for h in stalls}
    print(' ', r[0][-5:], r[S].rjust(5), r[E].rjust(8), r[1][:80], sorted(st.items(), key=lambda x: -x[1])[:2])
