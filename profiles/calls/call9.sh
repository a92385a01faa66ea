mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_gapt_n30_b512.csv python bench.py --workload train_gapt_n30_b512 --steps 2 --warmup 1 --no-graph --preload-s 0.0 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r2_launches_gapt_n30_b512.csv')))
# find header
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    try: v=float(r[mv].replace(',',''))
    except: continue
    name=r[kn].split('(')[0][-60:]
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]: print(f"{v[1]/1e3:10.1f} us {100*v[1]/tot:5.1f}%  n={v[0]:4d} avg {v[1]/v[0]/1e3:7.1f} us  {k}")
PY
