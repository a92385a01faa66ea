mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 8 --check > gpurun_out/r2_dp_check_8gpu.json 2> gpurun_out/r2_dp_check_8gpu.err; echo "check rc=$?"
python -c "import json; d=json.loads([l for l in open('gpurun_out/r2_dp_check_8gpu.json') if l.startswith('{')][-1]); print('dp check 8 gpus ok:', d['ok'], {k:{m:round(v[m]['sq_max_over_ranks'],6) for m in v} for k,v in d.items() if k.startswith('prec')})"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_suite_8gpu.json 2> gpurun_out/r2_bench_suite_8gpu.err; echo "suite rc=$?"
timeout 200 $TR profiles/gen_sweep.py 150 1000000 4096 2>/dev/null | tail -1 > gpurun_out/r2_gen_sweep_8gpu.txt
timeout 200 $TR profiles/gen_sweep.py 30 1000000 4096 2>/dev/null | tail -1 >> gpurun_out/r2_gen_sweep_8gpu.txt
cat gpurun_out/r2_gen_sweep_8gpu.txt
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_suite_8gpu.json') if l.startswith('{')][-1])
print('HEAD', round(d['value']), d['config'].get('collective'), d['ms_per_step'], d['e2e'])
for k,v in d.get('workloads',{}).items(): print(k, round(v.get('value',0)), round(v.get('ms_per_step',0),3), v.get('error'))
PY
