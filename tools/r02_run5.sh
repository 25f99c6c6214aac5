set -x
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -v > gpurun_out/r02e_gputests_full.txt 2>&1
grep -n "FAILED\|ERROR\|Fatal\|Segmentation\|passed\|failed" gpurun_out/r02e_gputests_full.txt | tail -20
timeout 300 python tools/bench_configs.py --configs c3,c5 --spp 16 > gpurun_out/r02e_configs_base.jsonl 2> gpurun_out/r02e_configs.err
for v in sb256x4 sb512x2m0 sb512x2m7 sb384x3; do
  GTLC_LIB_PATH=$PWD/gatling_b200/variants/libgtlc_$v.so timeout 300 python tools/bench_configs.py --configs c3,c5 --spp 16 > gpurun_out/r02e_configs_$v.jsonl 2>> gpurun_out/r02e_configs.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02e_configs_*.jsonl')):
    for l in open(f):
        d=json.loads(l)
        if d['sort']!=1: continue
        print(f.split('_')[-1][:-6], d['config'], {k:round(d[k],2) for k in ('Mray_s','ms_total','ms_trace','ms_shadow','ms_shade','ms_sort')})
PY
