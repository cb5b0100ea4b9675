timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "spot or asd or pairs or cfg3 or cfg4 or tiles or wide or 64" 2>&1 | tail -2
timeout 200 python scripts/fused_check.py 2>&1 | cut -c1-150
