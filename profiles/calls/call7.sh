mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR profiles/bench_peer.py 2>gpurun_out/r2_bench_peer.err | tee gpurun_out/r2_bench_peer_2gpu.txt
tail -3 gpurun_out/r2_bench_peer.err
for wl in train_n30_b256; do
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$wl 2gpu fused', round(d['value'],1), d['config'].get('collective'))"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --workload $wl --no-fused-allreduce 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$wl 2gpu nccl', round(d['value'],1), d['config'].get('collective'))"
done
