mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r2_pytest8.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest8.txt | tail -20
grep -n "^E  " gpurun_out/r2_pytest8.txt | head -20
for wl in train_gapt_n30_b512 train_gapt_isab_n30_b512; do
timeout 300 python bench.py --steps 20 --warmup 5 --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$wl', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'launches/step', d['gpu_launches']/d['steps'], 'hbm frac', round(d['roofline']['frac'],4), d['cpu_baseline'].get('gpu_eager'))"
done
