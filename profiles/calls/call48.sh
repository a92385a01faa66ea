timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
for wl in train_n30_b256 train_n30_b256_allreal train_n150_b32 gen_n30_b1024; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-suite --no-baselines --workload $wl 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['config']['workload'], round(l['value'],1), round(l['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in l['roofline'].get('kernels',{}).items()})"
done
