for v in "" _u; do echo "== lib$v"; GTLC_LIB_PATH=$PWD/gatling_b200/libgtlc$v.so timeout 600 python tools/bench_configs.py --configs c3,c5 --spp 16 --ups 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print(d['config'], d['sort'], round(d['Mray_s'],1), 'shade', round(d['ms_shade'],2), 'total', round(d['ms_total'],2))
"; done
