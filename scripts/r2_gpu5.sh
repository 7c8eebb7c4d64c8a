#!/bin/bash
# round 2, call 5: tests + default bench (with other_configs, copy-only leg) + launch list with instruction / smem counters
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest5.log 2>&1; echo "pytest: $(tail -n 1 gpurun_out/r2_pytest5.log)"
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench rc $?"; tail -n 4 gpurun_out/r2_bench5.err
python bench.py --steps 100 --no-other-configs --no-compaction --no-cpu-baseline --no-e2e > gpurun_out/r2_bench5_nocompact.json 2> gpurun_out/r2_bench5_nocompact.err; echo "bench(no compaction) rc $?"
C='python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 --no-other-configs'
$C > gpurun_out/plain5.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches5.csv $C > gpurun_out/ncu_launches5.log 2>&1
echo "ncu rc $?"; wc -l gpurun_out/r2_launches5.csv
