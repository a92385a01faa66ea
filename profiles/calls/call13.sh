mkdir -p gpurun_out
# launch lists (eager, cold-cache, serialised: shares only)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_n30_b256.csv python bench.py --workload train_n30_b256 --steps 2 --warmup 1 --no-graph --preload-s 0.0 --no-baselines > gpurun_out/r2_launches_n30.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_n150_b256.csv python bench.py --workload train_n150_b256 --steps 2 --warmup 1 --no-graph --preload-s 0.0 --no-baselines > gpurun_out/r2_launches_n150.log 2>&1
# full captures of the tcgen05 edge kernels: N=150 B=256 (p=0), N=30 B=256 with D's dropout
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_tc -c 3 -o gpurun_out/r2_edge_n150_b256 python profiles/run_edge.py 256 150 0.0 1 > gpurun_out/r2_ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_tc -c 3 -o gpurun_out/r2_edge_n30_b256_drop python profiles/run_edge.py 256 30 0.5 1 > gpurun_out/r2_ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
