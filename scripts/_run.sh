timeout 150 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | grep -E "Error|error|assert|passed|failed|differs" | head -20
