mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -80 > gpurun_out/r2_pytest5.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest5.txt | tail -10
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --check > gpurun_out/r2_dp_check_2gpu.json 2> gpurun_out/r2_dp_check_2gpu.err
echo "check rc=$?"; tail -c 2500 gpurun_out/r2_dp_check_2gpu.json; tail -5 gpurun_out/r2_dp_check_2gpu.err
date
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_suite_2gpu.json 2> gpurun_out/r2_bench_suite_2gpu.err
echo "suite rc=$?"; date
tail -3 gpurun_out/r2_bench_suite_2gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_suite_2gpu.json') if l.startswith('{')][-1])
print('HEAD', d['value'], d['config'].get('collective'))
for k,v in d.get('workloads',{}).items(): print(k, v.get('value'), v.get('error'))
PY
