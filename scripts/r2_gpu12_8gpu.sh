#!/bin/bash
# 8 GPUs: e2e vs copy-only (is the 1 -> 8 e2e curve the host side of the box or the library?)
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 100 --warmup 5 --ring 4 --no-other-configs --no-chained --profile-steps 0 --no-cpu-baseline > gpurun_out/r2_bench12_8gpu.json 2> gpurun_out/r2_bench12_8gpu.err; echo "bench8 rc $?"; tail -n 3 gpurun_out/r2_bench12_8gpu.err
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1; lscpu | head -20 >> gpurun_out/r2_topo.txt; numactl -H >> gpurun_out/r2_topo.txt 2>&1
