#!/bin/bash
# Round-2 evidence bundle on one B200: tests, both bench arms, ncu launch list with DRAM / instruction / shared-memory
# counters, one full capture of the dominant kernel, disagreement report.
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest: $(tail -n 1 gpurun_out/r2f_pytest.log)"
( time timeout 600 python bench.py --steps 200 --warmup 10 ) > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc $?"; tail -n 4 gpurun_out/r2f_bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc $?"
C='python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 --no-other-configs --no-chained'
$C > gpurun_out/r2f_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv $C > gpurun_out/r2f_ncu_launches.log 2>&1
echo "ncu launches rc $?"; wc -l gpurun_out/r2f_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lk_track -s 8 -c 1 -f -o gpurun_out/r2f_prof_lk $C > gpurun_out/r2f_ncu_full_lk.log 2>&1; echo "ncu full lk rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:select_kernel -s 8 -c 1 -f -o gpurun_out/r2f_prof_select $C > gpurun_out/r2f_ncu_full_select.log 2>&1; echo "ncu full select rc $?"
timeout 600 python scripts/disagreement.py --frames 200 --hd-frames 60 --out gpurun_out/r2_disagreement.json > gpurun_out/r2f_disagreement.log 2>&1; echo "disagreement rc $?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -3
