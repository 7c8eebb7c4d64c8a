#!/bin/bash
mkdir -p gpurun_out
cmd="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0"
RDFE_HARRIS_IMPL=3 RDFE_HARRIS_F2D=0 $cmd > gpurun_out/r2_plain4.log 2>&1 &&
RDFE_HARRIS_IMPL=3 RDFE_HARRIS_F2D=0 ncu --set full --import-source on --clock-control none -k regex:harris_nms3 -s 2 -c 1 -f -o gpurun_out/r2_prof_h3 $cmd > gpurun_out/r2_ncu4a.log 2>&1
RDFE_HARRIS_IMPL=0 ncu --set full --import-source on --clock-control none -k regex:harris_nms_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_h0 $cmd > gpurun_out/r2_ncu4b.log 2>&1
RDFE_HARRIS_IMPL=1 RDFE_HARRIS_MB=3 ncu --set full --import-source on --clock-control none -k regex:harris_flag -s 2 -c 1 -f -o gpurun_out/r2_prof_h1 $cmd > gpurun_out/r2_ncu4c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
