#!/bin/bash
mkdir -p gpurun_out
timeout 150 python scripts/lk_cache_ab.py 20 > gpurun_out/r2_lkab.log 2>&1; tail -n 3 gpurun_out/r2_lkab.log
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active -k regex:lk_track --clock-control none --csv --log-file gpurun_out/r2_lkab_ncu.csv python scripts/lk_cache_ab.py 2 > gpurun_out/r2_lkab_ncu.log 2>&1; echo "ncu rc $?"
timeout 120 python scripts/latency_breakdown.py 60 > gpurun_out/r2_latency_breakdown.txt 2>&1; cat gpurun_out/r2_latency_breakdown.txt
