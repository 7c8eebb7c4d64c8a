#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_edge_cases.py -q -k two_devices 2>&1 | tail -n 2
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc $?"; tail -n 3 gpurun_out/r2_bench_2gpu.err
