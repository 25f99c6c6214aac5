set -x
timeout 500 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
tail -c 400 gpurun_out/bench_c2.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/launches_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 8 -c 3 -f -o gpurun_out/trace_r01 python bench.py --steps 1 --warmup 1 --spp 8 --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
ls -la gpurun_out | tail -8
# host BVH8 builder on the box's cores (no GPU work): all threads and a quarter of them, with and without thread binding
timeout 600 python tools/bench_host_builder.py $(nproc) $(( $(nproc) / 4 )) > gpurun_out/host_builder.txt 2>&1
tail -30 gpurun_out/host_builder.txt
