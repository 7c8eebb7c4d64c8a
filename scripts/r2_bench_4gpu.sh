#!/bin/bash
mkdir -p gpurun_out
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err; echo "bench4 rc $?"; tail -n 5 gpurun_out/r2_bench_4gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29525 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 > gpurun_out/r2_bench_4gpu_ref.json 2> gpurun_out/r2_bench_4gpu_ref.err; echo "ref4 rc $?"
