mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -100 > gpurun_out/r2_pytest16.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest16.txt | tail -12
grep -n "^E  " gpurun_out/r2_pytest16.txt | head -12
python __graft_entry__.py --smoke 2>&1 | tail -3
