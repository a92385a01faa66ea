timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
MPG_LIB_VARIANT=trace timeout 120 python profiles/trace_fwd.py 512 30 0.5 2>&1 | tail -3 | cut -c1-60
MPG_LIB_VARIANT=trace timeout 120 python profiles/trace_chain.py 256 30 0.5 2>&1 | grep "mean period"
for wl in train_n30_b256 train_n150_b256; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-suite --no-baselines --workload $wl 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['config']['workload'], round(l['value'],1), round(l['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in l['roofline'].get('kernels',{}).items()})"
done
