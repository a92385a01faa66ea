export MPG_LIB_VARIANT=trace
for cfg in "1024 30 0.0" "512 30 0.5"; do
  echo "== $cfg"
  timeout 120 python profiles/trace_fwd.py $cfg 2>&1 | tail -5
done
