set -x
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -v -x > gpurun_out/r02c_gputests_full.txt 2>&1
grep -n "PASSED\|FAILED\|SKIPPED\|ERROR\|Fatal\|Segmentation" gpurun_out/r02c_gputests_full.txt | tail -50
grep -n "Current thread" -A25 gpurun_out/r02c_gputests_full.txt | head -60
