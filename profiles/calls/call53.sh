timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -12
