timeout 200 python profiles/trace_fn.py 15360 2>&1 | tail -50
