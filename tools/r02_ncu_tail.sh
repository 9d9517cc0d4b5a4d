#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 2 --no-cpu --no-profile"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/r02_launches_tail.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list rc $?"
ncu --set full --clock-control none --import-source on -k regex:k_mg_tail -s 6 -c 2 -o gpurun_out/r02_tail $CMD > gpurun_out/ncu_f.log 2>&1
echo "full rc $?"
tail -3 gpurun_out/ncu_f.log
