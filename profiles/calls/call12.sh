mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mab_.w._kernel -s 4 -c 2 -o gpurun_out/r2_mab_b512 python profiles/bench_mab.py 512 > gpurun_out/r2_mab_ncu.log 2>&1
tail -2 gpurun_out/r2_mab_ncu.log
