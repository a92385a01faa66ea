mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_n30_b256.csv python profiles/run_step.py 30 256 3 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_gapt_n30_b512.csv python profiles/run_step.py 30 512 3 gapt > /dev/null 2>&1
ls -la gpurun_out/r2_launches_n30_b256.csv gpurun_out/r2_launches_gapt_n30_b512.csv
