#!/bin/bash
# Round-end evidence, run on the GPU box (gpurun -- 'bash scripts/round_evidence.sh r02'): GPU test suite, both
# bench arms and the three other workloads, the ncu launch lists of the bench commands, one `ncu --set full` capture
# of every kernel at its BASELINE config size, compute-sanitizer over small shapes.  Everything lands in gpurun_out/;
# scripts/ncu_summary.py turns the .ncu-rep into profiles/ncu_full_<round>.md here (no GPU needed).
R=${1:-r02}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee gpurun_out/pytest_gpu_$R.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${R}_reference_arm.json 2> gpurun_out/bench_ref_$R.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${R}_n1.json 2> gpurun_out/bench_$R.err
tail -c 400 gpurun_out/bench_$R.err
for w in cfg2 cfg3 cfg4; do
  python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_${R}_$w.json 2> gpurun_out/bench_${R}_$w.err
  python bench.py --workload $w --impl reference --steps 1 --warmup 0 > gpurun_out/bench_${R}_${w}_reference_arm.json 2>> gpurun_out/bench_${R}_$w.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 3 --warmup 3 --min-seconds 0 --no-cpu --no-e2e --no-stages > gpurun_out/bench_under_ncu.log 2>&1
for w in cfg3 cfg4; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${R}_$w.csv \
      python bench.py --workload $w --steps 3 --warmup 3 --min-seconds 0 --no-cpu --no-e2e >> gpurun_out/bench_under_ncu.log 2>&1
done
PROFILE_REPS=1 ncu --set full --clock-control none \
    -k regex:"prep_kernel|simpool_kernel|rowreduce_kernel|topk_kernel|grouped_kernel|segmean_kernel|pair_cosine_kernel" \
    -o gpurun_out/prof_all_$R -f python scripts/profile_targets.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
for tool in memcheck synccheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_${tool}_$R.log python scripts/sanitize_target.py > gpurun_out/sanitizer_${tool}_$R.out 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_${tool}_$R.log
done
# racecheck once more with single-CTA MMAs: the only hazards the pair build reports sit inside tcgen05.alloc.cta_group::2
JEGAL_CTA_GROUP=1 timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck_cta1_$R.log python scripts/sanitize_target.py > gpurun_out/sanitizer_racecheck_cta1_$R.out 2>&1
tail -2 gpurun_out/sanitizer_racecheck_cta1_$R.log
python scripts/bench_grouped.py > gpurun_out/stages_$R.jsonl 2> gpurun_out/stages_$R.err
