mkdir -p gpurun_out
export MPG_LIB_VARIANT=trace
for cfg in "256 30 0.5" "512 30 0.5 rand" "256 30 0.0 rand"; do
  echo "== $cfg"
  timeout 120 python profiles/trace_chain.py $cfg 2>&1 | tail -9
done | tee gpurun_out/r2_trace_fixed.txt
