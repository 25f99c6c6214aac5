# A/B runs of build variants on the GPU box (how the round-2 sweeps in DESIGN.md section 3 were made).
#   1. build each variant HERE into gatling_b200/variants/ (git-ignored, travels with the gpurun snapshot), e.g.
#        nvcc ... -DGTLC_SHADE_BIG_BLOCK=1024 -o gatling_b200/variants/libgtlc_sb1024.so csrc/gtlc_api.cu csrc/bvh8_build.cpp
#      knobs: GTLC_SHADE_BIG_BLOCK, GTLC_SHADE_MIN_BLOCKS, GTLC_SHADE_PHASE_MASK, GTLC_TRACE_MIN_BLOCKS, GTLC_ENQUEUE_UNROLL
#   2. gpurun -- 'bash tools/r02_variants.sh sb1024 sb512x2 ...'
set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py --configs c3,c5 --spp 16 > gpurun_out/variants_base.jsonl 2> gpurun_out/variants.err
for v in "$@"; do
  GTLC_LIB_PATH=$PWD/gatling_b200/variants/libgtlc_$v.so timeout 300 python tools/bench_configs.py --configs c3,c5 --spp 16 > gpurun_out/variants_$v.jsonl 2>> gpurun_out/variants.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/variants_*.jsonl')):
    for l in open(f):
        d = json.loads(l)
        if d['sort'] == 1:
            print(f.split('variants_')[1][:-6], d['config'], {k: round(d[k], 2) for k in ('Mray_s', 'ms_total', 'ms_trace', 'ms_shadow', 'ms_shade', 'ms_sort')})
PY
