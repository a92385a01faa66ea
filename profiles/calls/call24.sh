mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:edge_tc --csv --log-file gpurun_out/r2_fixed_cost_drop.csv python profiles/fixed_cost.py 0.5 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:edge_tc --csv --log-file gpurun_out/r2_fixed_cost_nodrop.csv python profiles/fixed_cost.py 0.0 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ("gpurun_out/r2_fixed_cost_drop.csv", "gpurun_out/r2_fixed_cost_nodrop.csv"):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    print(f)
    for r in rows: print(r[4][:60], r[-1])
PY
