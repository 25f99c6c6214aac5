set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in A B C D; do
GTLC_LIB_PATH=$PWD/gatling_b200/variants/libgtlc_$v.so timeout 400 python tools/bench_configs.py --configs c3,c5 --spp 16 > gpurun_out/r02_configs_$v.jsonl 2> gpurun_out/r02_configs.err
python - <<PY
import json
for l in open('gpurun_out/r02_configs_$v.jsonl'):
    d=json.loads(l); print('$v', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('config','sort','Mray_s','ms_total','ms_trace','ms_shadow','ms_shade','ms_sort')})
PY
done
