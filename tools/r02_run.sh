set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/sweep_trace.py --spp 16 --variants 0,1 > gpurun_out/r02_sweep_c2.txt 2>&1
cat gpurun_out/r02_sweep_c2.txt
timeout 400 python tools/bench_configs.py --configs c3,c5 --spp 16 > gpurun_out/r02_configs.jsonl 2> gpurun_out/r02_configs.err
cat gpurun_out/r02_configs.jsonl; tail -3 gpurun_out/r02_configs.err
