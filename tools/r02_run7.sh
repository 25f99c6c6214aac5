set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py --configs c3,c5 --spp 16 > gpurun_out/r02g_configs_base.jsonl 2> gpurun_out/r02g_configs.err
GTLC_LIB_PATH=$PWD/gatling_b200/variants/libgtlc_eq1.so timeout 300 python tools/bench_configs.py --configs c3,c5 --spp 16 > gpurun_out/r02g_configs_eq1.jsonl 2>> gpurun_out/r02g_configs.err
timeout 300 python tools/sweep_trace.py --spp 16 --variants 0 --refill 4,6,8,12,16 --ctas 6,7,8 2>&1 | tail -20 | tee gpurun_out/r02g_sweep_c2.txt
GTLC_LIB_PATH=$PWD/gatling_b200/variants/libgtlc_eq1.so timeout 300 python tools/sweep_trace.py --spp 16 --variants 0 --refill 8 --ctas 8 2>&1 | tail -3 | tee gpurun_out/r02g_sweep_c2_eq1.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02g_configs_*.jsonl')):
    for l in open(f):
        d=json.loads(l)
        if d['sort']!=1: continue
        print(f.split('_')[-1][:-6], d['config'], {k:round(d[k],2) for k in ('Mray_s','ms_total','ms_trace','ms_shadow','ms_shade','ms_sort')})
PY
