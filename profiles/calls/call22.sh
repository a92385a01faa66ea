mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -100 > gpurun_out/r2_pytest22.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest22.txt | tail -12
grep -n "^E  " gpurun_out/r2_pytest22.txt | head -12
python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 200 python profiles/bench_mab.py 2>&1 | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_suite_1gpu.json 2> gpurun_out/r2_bench22.err
tail -3 gpurun_out/r2_bench22.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_suite_1gpu.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('metric','value','ms_per_step','gpu_launches')}, l.get('e2e'), l.get('roofline'))
for k,v in l.get('workloads',{}).items(): print(k, round(v.get('value',0),1), v.get('ms_per_step'), v.get('gpu_launches'))
PY
