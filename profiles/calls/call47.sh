mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_suite_1gpu.json 2> gpurun_out/r2_bench47.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_suite_1gpu.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('metric','value','ms_per_step','gpu_launches')}, l.get('e2e'), l['clocks'])
r=l['roofline']; print({k:r[k] for k in ('kernel','achieved','frac','traffic','share_of_step','executed_step_fraction')})
for k,v in l.get('workloads',{}).items(): print(k, round(v.get('value',0),1), round(v.get('ms_per_step',0),4), v.get('gpu_launches'), round(v.get('roofline',{}).get('frac',0),3), v.get('roofline',{}).get('executed_step_fraction'))
PY
