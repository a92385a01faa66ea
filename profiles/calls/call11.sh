for B in 37 148 296 592 1184; do timeout 100 python profiles/bench_mab.py $B 2>&1 | grep "Nq=30 Nk=30 fused"; done
