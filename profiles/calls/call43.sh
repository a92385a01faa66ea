mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_tc -c 3 -o gpurun_out/r2_edge_n150_b256 python profiles/run_edge.py 256 150 0.0 1 > gpurun_out/r2_ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_tc -c 3 -o gpurun_out/r2_edge_n30_b256_drop python profiles/run_edge.py 256 30 0.5 1 > gpurun_out/r2_ncu_b.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_n30_b256.csv python profiles/run_step.py 30 256 3 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_n150_b256.csv python profiles/run_step.py 150 256 2 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_gapt_n30_b512.csv python profiles/run_step.py 30 512 3 gapt > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_*.csv
