mkdir -p gpurun_out
timeout 200 python profiles/bench_mab.py 512 2>&1 | tail -10
