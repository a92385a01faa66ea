mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_features.py tests/test_gpu_parity.py -m gpu -q -k "mab or gapt" 2>&1 | tail -40 > gpurun_out/r2_pytest20.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest20.txt | tail -14
grep -n "^E  " gpurun_out/r2_pytest20.txt | head -10
timeout 200 python profiles/bench_mab.py 512 2>&1 | grep fused
for wl in train_gapt_n30_b512 train_gapt_isab_n30_b512; do
timeout 300 python bench.py --steps 20 --warmup 5 --workload $wl --no-baselines 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$wl', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'launches/step', d['gpu_launches']/d['steps'])"
done
