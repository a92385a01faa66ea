mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -120 > gpurun_out/r2_pytest6.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest6.txt | tail -10
grep -n "^E  " gpurun_out/r2_pytest6.txt | head -10
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --check 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('dp check ok:', d['ok'])"
for wl in train_n30_b256 train_n150_b256; do
timeout 300 python bench.py --steps 20 --warmup 5 --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$wl 1gpu', round(d['value'],1))"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$wl 2gpu fused', round(d['value'],1), d['config'].get('collective'))"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --workload $wl --no-fused-allreduce 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$wl 2gpu nccl', round(d['value'],1), d['config'].get('collective'))"
done
