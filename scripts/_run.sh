mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 scripts/h2d_ceiling.py > gpurun_out/h2d_ceiling_n8.json 2> gpurun_out/h2d_n8.err; tail -c 300 gpurun_out/h2d_n8.err; cat gpurun_out/h2d_ceiling_n8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r02_n8.json 2> gpurun_out/bench_r02_n8.err; tail -c 500 gpurun_out/bench_r02_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_r02_n4.json 2> gpurun_out/bench_r02_n4.err; tail -c 500 gpurun_out/bench_r02_n4.err
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
