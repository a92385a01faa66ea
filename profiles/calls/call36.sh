timeout 200 python profiles/trace_fn.py 15360 2>&1 | grep -E "fwd|bwd|A tile|exit"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "node_net" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-suite --no-baselines --workload train_n30_b256 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['config']['workload'], round(l['value'],1), round(l['ms_per_step'],4))"
