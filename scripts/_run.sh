cd /root/repo
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_r02_n4.json 2> gpurun_out/bench_r02_n4.err; echo "rc=$?"; tail -c 400 gpurun_out/bench_r02_n4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02_n4.json").read().strip().splitlines()[-1])
print(d["steps"], d["ms_per_step"], d["value"], d["sustained"], d["clocks"], d["e2e"]["ms_per_step"], d["e2e"]["steps"], d["gpu_launches"], d["roofline"]["frac"], d["config"]["topk_checksum"], d["config"]["recall_at_1_planted"])
PY
