set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_pytest1.txt
timeout 300 python profiles/error_table.py > gpurun_out/r2_error_table.txt 2>&1
MPG_LIB_VARIANT=trace timeout 120 python profiles/trace_chain.py 256 150 0.0 > gpurun_out/r2_trace_chain_n150.txt 2>&1
MPG_LIB_VARIANT=trace timeout 120 python profiles/trace_chain.py 256 30 0.5 > gpurun_out/r2_trace_chain_n30.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_suite_a.json 2> gpurun_out/r2_bench_suite_a.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_tc -c 3 -o gpurun_out/r2a_edge_n150_b256 python profiles/run_edge.py 256 150 0.0 1 > gpurun_out/r2a_ncu.log 2>&1
tail -5 gpurun_out/r2_pytest1.txt
tail -3 gpurun_out/r2_error_table.txt
cat gpurun_out/r2_trace_chain_n150.txt
