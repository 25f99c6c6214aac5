set -x
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -v > gpurun_out/r02j_gputests_full.txt 2>&1
grep -n "FAILED\|ERROR\|Fatal\|passed\|failed" gpurun_out/r02j_gputests_full.txt | tail -12
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
