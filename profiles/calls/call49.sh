timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_n30_b256.csv python profiles/run_step.py 30 256 3 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_n30_b256.csv')) if len(r)>10 and r[0].isdigit()]
n=len(rows)//3; last=rows[-n:]
agg=collections.OrderedDict()
for r in last:
    k=r[4].split('(')[0][-44:]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[-1])/1e3
tot=sum(v[1] for v in agg.values()); print(len(last), round(tot,1))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]: print(f"   {v[1]:9.1f} us {v[0]:3d}x {k}")
PY
