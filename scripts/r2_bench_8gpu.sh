#!/bin/bash
mkdir -p gpurun_out
( time timeout 560 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 ) > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "bench8 rc $?"; tail -n 5 gpurun_out/r2_bench_8gpu.err
