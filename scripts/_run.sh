mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 --e2e-timeline > gpurun_out/bench_r02_n8.json 2> gpurun_out/bench_r02_n8.err; grep -o '{"rank".*' gpurun_out/bench_r02_n8.err > gpurun_out/e2e_timeline_n8.jsonl
