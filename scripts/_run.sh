cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scripts.py -x -q -m gpu -k "spot or asd or pairs or k3 or K3 or stream or script or wide or golden" 2>&1 | tail -4
timeout 200 python scripts/fused_check.py 2>&1 | tail -12
JEGAL_B200_LIB=jegal_b200/libjegal_b200_trace.so timeout 200 python scripts/grouped_trace.py 2>&1 | tail -3
