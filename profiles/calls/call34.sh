for i in 1 2 3 4 5 6 7 8 9 10; do
timeout 300 python -m pytest tests/test_gpu_train_mode.py -m gpu -q -x -k "graph_replay" 2>&1 | grep -E "^E  |passed|failed" | head -6
done
