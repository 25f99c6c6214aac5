# Round-2 GPU session: tests, both bench arms, launch list, DRAM traffic of one whole wave per config, ncu --set full captures of the
# traversal and shade kernels on C3 (and the traversal kernel on C2).  gpurun copies back at most 64 MiB: the reports are reduced to
# CSV / text on the box and only kept while the directory stays under that.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
if [ -z "$SKIP_TESTS" ]; then timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_gputests.txt; fi
timeout 1200 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 300 gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
timeout 300 python tools/bench_configs.py --configs c1,c3,c5 --spp 16 > gpurun_out/r02_configs.jsonl 2> gpurun_out/r02_configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/r02_launches_run.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k 'regex:trace_kernel_batched|shade_kernel' -s 128 -c 24 --csv --log-file gpurun_out/r02_traffic_c3.csv python tools/bench_configs.py --configs c3 --spp 8 > gpurun_out/r02_traffic_c3.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k 'regex:trace_kernel_batched|shade_kernel' -s 128 -c 24 --csv --log-file gpurun_out/r02_traffic_c5.csv python tools/bench_configs.py --configs c5 --spp 8 > gpurun_out/r02_traffic_c5.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k 'regex:trace_kernel_batched|shade_kernel' -s 56 -c 16 --csv --log-file gpurun_out/r02_traffic_c2.csv python bench.py --config c2 --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/r02_traffic_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel_batched -s 48 -c 4 -f -o gpurun_out/r02_trace_c3 python tools/bench_configs.py --configs c3 --spp 8 > gpurun_out/r02_ncu_trace_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shade_kernel -s 32 -c 2 -f -o gpurun_out/r02_shade_c3 python tools/bench_configs.py --configs c3 --spp 8 > gpurun_out/r02_ncu_shade_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel_batched -s 24 -c 2 -f -o gpurun_out/r02_trace_c2 python bench.py --config c2 --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/r02_ncu_trace_c2.log 2>&1
for r in trace_c3 shade_c3 trace_c2; do
  ncu -i gpurun_out/r02_$r.ncu-rep --page raw --csv > gpurun_out/r02_$r.raw.csv 2>/dev/null
  python tools/ncu_phases.py gpurun_out/r02_$r.ncu-rep --launch 0 --top 40 > gpurun_out/r02_$r.lines0.txt 2>&1
  python tools/ncu_phases.py gpurun_out/r02_$r.ncu-rep --launch 1 --top 40 > gpurun_out/r02_$r.lines1.txt 2>&1
done
python tools/ncu_phases.py gpurun_out/r02_trace_c3.ncu-rep --launch 2 --top 10 --ranges "refill:traverse.cuh:717-752,nodestep:traverse.cuh:753-759,enqueue:traverse.cuh:760-771,drain:traverse.cuh:772-804,finish:traverse.cuh:805-832,nodefn:traverse.cuh:262-318,trifn:traverse.cuh:223-261,setup:traverse.cuh:145-172" > gpurun_out/r02_trace_c3.phases2.txt 2>&1
du -sm gpurun_out; ls -la gpurun_out
for r in trace_c2 shade_c3 trace_c3; do
  if [ $(du -sm gpurun_out | cut -f1) -gt 56 ]; then rm -f gpurun_out/r02_$r.ncu-rep; fi
done
du -sm gpurun_out
