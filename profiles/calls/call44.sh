timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
for wl in train_n30_b256 train_n150_b32 gen_n30_b1024; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-suite --no-baselines --workload $wl 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['config']['workload'], round(l['value'],1), round(l['ms_per_step'],4))"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_n30_b256.csv python profiles/run_step.py 30 256 3 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_n30_b256.csv')) if len(r)>10 and r[0].isdigit()]
n=len(rows)//3; last=rows[-n:]
agg=collections.OrderedDict()
for r in last:
    k=r[4].split('(')[0][-44:]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[-1])/1e3
tot=sum(v[1] for v in agg.values()); print(len(last), round(tot,1))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f"   {v[1]:9.1f} us {v[0]:3d}x {k}")
PY
