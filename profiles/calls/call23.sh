mkdir -p gpurun_out
timeout 120 ./profiles/microbench/pdl_bench 2>&1 | tee gpurun_out/r2_pdl_bench.txt
timeout 300 python -m pytest tests/test_gpu_features.py -m gpu -q -k "wgan_gp" 2>&1 | tail -5
