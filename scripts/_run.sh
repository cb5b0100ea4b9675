cd /root/repo
mkdir -p gpurun_out
R=r02
for tool in racecheck memcheck synccheck; do
  timeout 110 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_${tool}_$R.log python scripts/sanitize_target.py > gpurun_out/sanitizer_${tool}_$R.out 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/sanitizer_${tool}_$R.log
done
