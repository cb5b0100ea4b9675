timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide or 64 or spot or asd" 2>&1 | tail -4
