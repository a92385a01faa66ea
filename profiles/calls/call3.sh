mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -60 > gpurun_out/r2_pytest3.txt
tail -4 gpurun_out/r2_pytest3.txt
MPG_LIB_VARIANT=trace timeout 120 python profiles/trace_chain.py 256 150 0.0 2>&1 | tail -4
MPG_LIB_VARIANT=trace timeout 120 python profiles/trace_chain.py 256 30 0.5 2>&1 | tail -2
echo "== product lib"; timeout 120 python profiles/run_edge.py 256 150 0.0 4 | tail -2
echo "== test_wait variant"; MPG_LIB_VARIANT=testwait timeout 120 python profiles/run_edge.py 256 150 0.0 4 | tail -2
for wl in train_n30_b256 train_n150_b256_allreal; do
for v in "" testwait; do
MPG_LIB_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 5 --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl','$v', round(d['value'],1), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done; done
