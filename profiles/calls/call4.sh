mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r2_pytest4.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest4.txt | tail -30
