#!/bin/bash
# Round-end evidence, run on the GPU box (gpurun -- 'bash scripts/round_evidence.sh r01'): GPU test suite, both
# bench arms, stage timings, CPU-vs-GPU stage table, the ncu launch list of the bench command and one
# `ncu --set full` capture of every kernel at its BASELINE config size.  Everything lands in gpurun_out/;
# scripts/ncu_summary.py turns the .ncu-rep into profiles/ncu_full_<round>.md here (no GPU needed).
R=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -c 400 gpurun_out/bench_$R.err
python scripts/bench_grouped.py > gpurun_out/stages_$R.jsonl 2> gpurun_out/stages_$R.err
python scripts/cpu_vs_gpu_stages.py > gpurun_out/cpu_vs_gpu_$R.jsonl 2> gpurun_out/cpu_vs_gpu_$R.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
PROFILE_REPS=1 ncu --set full --clock-control none \
    -k regex:"prep_kernel|simpool_kernel|rowreduce_kernel|topk_kernel|grouped_kernel|segmean_kernel" \
    -o gpurun_out/prof_all_$R -f python scripts/profile_targets.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
