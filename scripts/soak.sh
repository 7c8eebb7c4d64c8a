#!/bin/bash
# GPU box: lock-step GPU-vs-oracle soak over several workloads in parallel (tests/diag_replay.py)
D="python tests/diag_replay.py"
$D --stream 3 --frames 200 > gpurun_out/d0.log 2>&1 &
$D --stream 11 --frames 200 --pred noisy --cache > gpurun_out/d1.log 2>&1 &
$D --stream 12 --frames 150 --pred bad > gpurun_out/d2.log 2>&1 &
$D --stream 13 --frames 66 --skip 3 --cache > gpurun_out/d3.log 2>&1 &
$D --workload advio --stream 5 --frames 60 --cache > gpurun_out/d4.log 2>&1 &
$D --workload advio --stream 6 --frames 60 --pred noisy > gpurun_out/d5.log 2>&1 &
$D --workload hd --stream 7 --frames 30 --cache > gpurun_out/d6.log 2>&1 &
$D --workload hd --stream 8 --frames 30 --pred noisy > gpurun_out/d7.log 2>&1 &
$D --workload hd --stream 9 --frames 20 --pred bad --cache > gpurun_out/d8.log 2>&1 &
wait
tail -q -n 1 gpurun_out/d?.log
