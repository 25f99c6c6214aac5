# Round-2 GPU session: both bench arms, launch list, ncu --set full captures of the traversal and shade kernels on C3 (and the
# traversal kernel on C2).  gpurun copies back at most 64 MiB: the reports are reduced to CSV / text on the box and only kept
# while the directory stays under that.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
if [ -z "$SKIP_TESTS" ]; then timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_gputests.txt; fi
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 300 gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
timeout 300 python tools/bench_configs.py --configs c3,c5 --spp 16 > gpurun_out/r02_configs.jsonl 2> gpurun_out/r02_configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/r02_launches_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel_batched -s 48 -c 4 -f -o gpurun_out/r02_trace_c3 python tools/bench_configs.py --configs c3 --spp 4 > gpurun_out/r02_ncu_trace_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shade_kernel -s 24 -c 2 -f -o gpurun_out/r02_shade_c3 python tools/bench_configs.py --configs c3 --spp 4 > gpurun_out/r02_ncu_shade_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel_batched -s 8 -c 2 -f -o gpurun_out/r02_trace_c2 python bench.py --config c2 --steps 1 --warmup 1 --spp 8 --no-cpu-baseline > gpurun_out/r02_ncu_trace_c2.log 2>&1
for r in trace_c3 shade_c3 trace_c2; do
  ncu -i gpurun_out/r02_$r.ncu-rep --page raw --csv > gpurun_out/r02_$r.raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_$r.ncu-rep --page details --csv > gpurun_out/r02_$r.details.csv 2>/dev/null
  python tools/ncu_phases.py gpurun_out/r02_$r.ncu-rep --launch 0 --top 40 > gpurun_out/r02_$r.lines0.txt 2>&1
  python tools/ncu_phases.py gpurun_out/r02_$r.ncu-rep --launch 1 --top 40 > gpurun_out/r02_$r.lines1.txt 2>&1
done
ncu -i gpurun_out/r02_shade_c3.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_shade_c3.sass.csv 2>/dev/null; gzip -f gpurun_out/r02_shade_c3.sass.csv
du -sm gpurun_out; ls -la gpurun_out
for r in trace_c2 trace_c3 shade_c3; do
  if [ $(du -sm gpurun_out | cut -f1) -gt 56 ]; then rm -f gpurun_out/r02_$r.ncu-rep; fi
done
du -sm gpurun_out
