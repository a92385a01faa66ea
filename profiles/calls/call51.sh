timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
for wl in gen_n30_b1024 gen_n150_b1024 train_n30_b256; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-suite --no-baselines --workload $wl 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['config']['workload'], round(l['value'],1), round(l['ms_per_step'],4))"
done
