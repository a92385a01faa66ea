export MPG_LIB_VARIANT=trace
for cfg in "256 30 0.5" "256 30 0.0 rand"; do
  echo "== $cfg"
  timeout 120 python profiles/trace_chain.py $cfg 2>&1 | tail -4
done
