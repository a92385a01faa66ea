mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mab_fwd_tc_kernel -s 4 -c 1 -o gpurun_out/r2_mab_tc_fwd python profiles/bench_mab.py 512 > gpurun_out/r2_mab_tc_ncu.log 2>&1
tail -1 gpurun_out/r2_mab_tc_ncu.log
