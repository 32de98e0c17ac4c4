# compute-sanitizer memcheck (and racecheck on the env / tree kernels) over small workloads that touch every kernel family; logs go to
# gpurun_out/<tag>_san_*.log, the summary lines are copied into profiles/ by hand.   bash tools/sanitize_all.sh r2
TAG=${1:-r2}
for t in sanitize_env sanitize_train; do
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/$t.py > gpurun_out/${TAG}_san_memcheck_$t.log 2>&1
  echo "memcheck $t rc=$?" | tee -a gpurun_out/${TAG}_san_summary.txt
  grep -E "ERROR SUMMARY|Invalid|Out of bounds" gpurun_out/${TAG}_san_memcheck_$t.log | tail -3 | tee -a gpurun_out/${TAG}_san_summary.txt
done
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_env.py > gpurun_out/${TAG}_san_racecheck_env.log 2>&1
echo "racecheck sanitize_env rc=$?" | tee -a gpurun_out/${TAG}_san_summary.txt
grep -E "RACECHECK SUMMARY|hazard" gpurun_out/${TAG}_san_racecheck_env.log | tail -3 | tee -a gpurun_out/${TAG}_san_summary.txt
