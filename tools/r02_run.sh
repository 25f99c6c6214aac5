set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 400 python tools/bench_configs.py --configs c3,c5 --spp 16 > gpurun_out/r02_configs.jsonl 2> gpurun_out/r02_configs.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_configs.jsonl'):
    d=json.loads(l); print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('config','sort','Mray_s','ms_total','ms_trace','ms_shadow','ms_shade','ms_sort','launches')})
PY
tail -3 gpurun_out/r02_configs.err
