set -x
mkdir -p gpurun_out
GTLC_LIB_PATH=$PWD/gatling_b200/variants/libgtlc_$1.so timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r02_gputests_$1.txt
bash tools/r02_variants.sh "$@" 2>&1 | tail -8
