#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:poisson_append -s 30 -c 1 -f -o gpurun_out/r2_prof_poisson1 python scripts/latency_breakdown.py 45 > gpurun_out/r2_ncu_p1.log 2>&1; echo "rc $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:select_kernel -s 30 -c 1 -f -o gpurun_out/r2_prof_select1 python scripts/latency_breakdown.py 45 > gpurun_out/r2_ncu_s1.log 2>&1; echo "rc $?"
ls -la gpurun_out/*1.ncu-rep
