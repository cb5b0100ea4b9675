cd /root/repo
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02.err; echo "rc=$?"; tail -c 300 gpurun_out/bench_r02.err
timeout 200 python bench.py --workload cfg3 --steps 20 --warmup 3 > gpurun_out/bench_r02_cfg3.json 2> gpurun_out/bench_r02_cfg3.err; echo "rc=$?"; tail -c 300 gpurun_out/bench_r02_cfg3.err
python - <<'PY'
import json
for f in ["gpurun_out/bench_r02_n1.json","gpurun_out/bench_r02_cfg3.json"]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["steps"], d["ms_per_step"], d["sustained"], d["clocks"], d["e2e"]["ms_per_step"], d["e2e"]["steps"], d["gpu_launches"], d["roofline"]["frac"])
PY
