MPG_LIB_VARIANT=trace timeout 120 python profiles/trace_fwd.py 512 30 0.5 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "dropout or edge_tc" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-suite --no-baselines --workload train_n30_b256 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['config']['workload'], round(l['value'],1), round(l['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in l['roofline'].get('kernels',{}).items()})"
