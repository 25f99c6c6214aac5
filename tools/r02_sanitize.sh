# compute-sanitizer pass over every kernel family (tools/sanitize_check.py):  gpurun -- 'bash tools/r02_sanitize.sh'
set -x
mkdir -p gpurun_out
timeout 300 python tools/sanitize_check.py 2>&1 | tail -2 | tee gpurun_out/r02_sanitize_plain.txt
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_check.py > gpurun_out/r02_sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r02_sanitize_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE PASS" gpurun_out/r02_sanitize_$tool.log | tee -a gpurun_out/r02_sanitize_summary.txt
  tail -c 3000 gpurun_out/r02_sanitize_$tool.log > gpurun_out/r02_sanitize_$tool.tail.txt
  head -c 20000 gpurun_out/r02_sanitize_$tool.log > gpurun_out/r02_sanitize_$tool.head.txt
  rm -f gpurun_out/r02_sanitize_$tool.log
done
