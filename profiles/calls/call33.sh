timeout 600 python -m pytest tests/test_gpu_train_mode.py -m gpu -q -x -k "graph_replay" 2>&1 | grep -E "^E |assert|Error|passed|failed" | head -30
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6
