set -x
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -v > gpurun_out/r02d_gputests_full.txt 2>&1
grep -n "PASSED\|FAILED\|SKIPPED\|ERROR\|Fatal\|Segmentation\|passed\|failed" gpurun_out/r02d_gputests_full.txt | tail -60
timeout 300 python tools/bench_configs.py --configs c3,c5 --spp 16 > gpurun_out/r02d_configs_base.jsonl 2> gpurun_out/r02d_configs.err
for v in sb896 sb1024 sb512x2; do
  GTLC_LIB_PATH=$PWD/gatling_b200/variants/libgtlc_$v.so timeout 300 python tools/bench_configs.py --configs c3,c5 --spp 16 > gpurun_out/r02d_configs_$v.jsonl 2>> gpurun_out/r02d_configs.err
done
timeout 300 python tools/bench_configs.py --configs c3,c5 --spp 16 --bvh-builder 2 > gpurun_out/r02d_configs_ploc.jsonl 2>> gpurun_out/r02d_configs.err
timeout 300 python tools/bench_configs.py --configs c3 --spp 16 --trace-variant 6 > gpurun_out/r02d_configs_tma.jsonl 2>> gpurun_out/r02d_configs.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02d_configs_*.jsonl')):
    for l in open(f):
        d=json.loads(l)
        if d['sort']!=1: continue
        print(f.split('_')[-1][:-6], d['config'], {k:round(d[k],2) for k in ('Mray_s','ms_total','ms_trace','ms_shadow','ms_shade','ms_sort','bvh_build_ms')})
PY
