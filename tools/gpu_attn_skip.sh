for sk in 0 1 2 3 4 8 15; do AF2_ATTN_SKIP=$sk AF2_ATTN_TRACE=1 timeout 100 python tools/attn_trace.py 256 > gpurun_out/attn_skip_${sk}.txt 2>&1; python - $sk <<'PY'
import sys,re
sk=sys.argv[1]
per=[]
for l in open(f'gpurun_out/attn_skip_{sk}.txt'):
    m=re.search(r'period (\d+)',l)
    if m: per.append(int(m.group(1)))
p=per[10:50]
print('skip mask',sk,'mean period per block',sum(p)/max(len(p),1))
PY
done
