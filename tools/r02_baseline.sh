set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/bench_configs.py --configs c3,c5 --spp 16 > gpurun_out/r02_base_configs.jsonl 2> gpurun_out/r02_base_configs.err
cat gpurun_out/r02_base_configs.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shade_kernel -s 24 -c 3 -f -o gpurun_out/r02_shade_c3_base python tools/bench_configs.py --configs c3 --spp 4 > gpurun_out/r02_ncu_shade.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel_batched -s 48 -c 4 -f -o gpurun_out/r02_trace_c3_base python tools/bench_configs.py --configs c3 --spp 4 > gpurun_out/r02_ncu_trace.log 2>&1
ls -la gpurun_out | tail -5
