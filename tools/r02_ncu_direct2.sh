#!/bin/bash
mkdir -p gpurun_out
for w in pillow pringle monkey_saddle dome arch; do
  CMD="python bench.py --workload $w --steps 3 --warmup 2 --no-cpu"
  $CMD > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum --clock-control none -k regex:k_direct --csv --log-file gpurun_out/r02_direct_$w.csv $CMD > gpurun_out/ncu_l.log 2>&1
  echo "$w rc $?"
done
CMD="python bench.py --workload dome --steps 3 --warmup 2 --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:k_direct_warp -s 2 -c 1 -o gpurun_out/r02_direct_warp2 $CMD > gpurun_out/ncu_f.log 2>&1
echo "full rc $?"
