mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; tail -c 300 gpurun_out/bench_r02_n1.err
