set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -v 2>&1 | tail -8 | tee gpurun_out/r02f_multigpu_test.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 --sub-configs c2 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err
tail -c 1500 gpurun_out/r02f_bench_n2.json; tail -5 gpurun_out/r02f_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02f_bench_ref_n2.json 2> gpurun_out/r02f_bench_ref_n2.err
tail -c 600 gpurun_out/r02f_bench_ref_n2.json
