mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
MPG_MULTICAST_MIN_WORLD=2 timeout 300 $TR bench.py --gpus 2 --check 2>gpurun_out/c15.err | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('dp check (multimem) ok:', d['ok'], d['precision0']['fused'])"
tail -3 gpurun_out/c15.err
timeout 200 $TR profiles/bench_peer.py 2>/dev/null | tee gpurun_out/r2_bench_peer_2gpu.txt
