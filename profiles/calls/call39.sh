mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 8 --check > gpurun_out/r2_dp_check_8gpu.json 2> gpurun_out/r2_dp_check_8gpu.err; echo "check rc=$?"
python -c "import json; d=json.loads([l for l in open('gpurun_out/r2_dp_check_8gpu.json') if l.startswith('{')][-1]); print('dp check 8 gpus ok:', d['ok'], {k:{m:round(v[m]['sq_max_over_ranks'],6) for m in v} for k,v in d.items() if k.startswith('prec')})"
timeout 200 $TR profiles/bench_peer.py 2>/dev/null | tee gpurun_out/r2_bench_peer_8gpu.txt
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_suite_8gpu.json 2> gpurun_out/r2_bench_suite_8gpu.err; echo "suite rc=$?"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 --workload train_n30_b256 --no-suite --no-baselines --no-fused-allreduce > gpurun_out/r2_bench_n30_8gpu_nccl.json 2>/dev/null
MPG_MULTICAST_MIN_WORLD=99 timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 --workload train_n30_b256 --no-suite --no-baselines > gpurun_out/r2_bench_n30_8gpu_peerloads.json 2>/dev/null
timeout 300 $TR profiles/gen_sweep.py 150 1000000 4096 2>/dev/null | tail -1 > gpurun_out/r2_gen_sweep_8gpu.txt
timeout 300 $TR profiles/gen_sweep.py 30 1000000 4096 2>/dev/null | tail -1 >> gpurun_out/r2_gen_sweep_8gpu.txt
cat gpurun_out/r2_gen_sweep_8gpu.txt
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_suite_8gpu.json') if l.startswith('{')][-1])
print('HEAD', round(d['value']), d['config'].get('collective'), d['ms_per_step'])
for k,v in d.get('workloads',{}).items(): print(k, round(v.get('value',0)), v.get('error'))
for f in ('nccl','peerloads'):
    n=json.loads([l for l in open(f'gpurun_out/r2_bench_n30_8gpu_{f}.json') if l.startswith('{')][-1]); print(f, round(n['value']), n['config'].get('collective'))
PY
