set -x
mkdir -p gpurun_out
GTLC_LIB_PATH=$PWD/gatling_b200/variants/libgtlc_aos.so timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r02i_gputests_aos.txt
bash tools/r02_variants.sh hit32 aos 2>&1 | tail -8
