#!/bin/bash
# GPU box, one GPU: the evidence bundle of a round (tests, both bench arms, ncu launch list with DRAM bytes).
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest: $(tail -n 1 gpurun_out/pytest.log)"
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc $?"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc $?"
python bench.py --workload advio --streams 32 --no-cpu-baseline > gpurun_out/bench_advio.json 2> gpurun_out/bench_advio.err; echo "advio rc $?"
python bench.py --workload hd --streams 32 --no-cpu-baseline > gpurun_out/bench_hd.json 2> gpurun_out/bench_hd.err; echo "hd rc $?"
C='python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0'
$C > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv $C > gpurun_out/ncu_launches.log 2>&1
echo "ncu rc $?"; wc -l gpurun_out/launches_final.csv
# one --set full capture of the kernel changed last (harris_nms after the narrow-tile change), steady state
timeout 120 ncu --set full --clock-control none --import-source on -k regex:harris_nms -s 12 -c 1 -f -o gpurun_out/prof_r1_harris_narrow $C > gpurun_out/ncu_full_harris.log 2>&1
echo "ncu full rc $?"
