# Round-2 check of sample-range sharding on an N-GPU box:  gpurun --gpus N -- 'bash tools/r02_sample_shard.sh N'
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02_gputests_sampleshard.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 tools/multigpu_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tee gpurun_out/r02_multigpu_check_samples_n$N.txt
for mode in tiles samples; do
  timeout 900 $TR --master-port 29612 bench.py --gpus $N --steps 3 --warmup 3 --sub-configs '' --no-cpu-baseline --shard $mode > gpurun_out/r02_bench_n${N}_$mode.json 2> gpurun_out/r02_bench_n${N}_$mode.err
  tail -c 600 gpurun_out/r02_bench_n${N}_$mode.json | head -c 600; echo
done
python - <<PY
import json
for m in ("tiles", "samples"):
    d = json.loads(open("gpurun_out/r02_bench_n${N}_%s.json" % m).read().strip().splitlines()[-1])
    print(m, d["value"], d["unit"], d["ms_per_step"], d.get("e2e", {}).get("value"))
PY
