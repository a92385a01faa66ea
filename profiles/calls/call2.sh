mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -60 > gpurun_out/r2_pytest2.txt
tail -5 gpurun_out/r2_pytest2.txt
MPG_LIB_VARIANT=trace timeout 120 python profiles/trace_chain.py 256 150 0.0 2>&1 | tail -4
MPG_LIB_VARIANT=trace timeout 120 python profiles/trace_chain.py 256 30 0.5 2>&1 | tail -2
timeout 120 python profiles/run_edge.py 256 150 0.0 3
timeout 300 python bench.py --steps 20 --warmup 5 --workload train_n30_b256 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n30', d['value'], d['roofline']['kernels'])"
timeout 300 python bench.py --steps 20 --warmup 5 --workload train_n150_b256_allreal 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n150ar', d['value'], d['roofline']['kernels'])"
