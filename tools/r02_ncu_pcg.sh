#!/bin/bash
# round-2 evidence for the PCG path: launch list of one whole step, then --set full on one iteration's kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --no-cpu --no-profile"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 2700 -c 1000 --csv --log-file gpurun_out/r02_grid256_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list rc $?"
ncu --set full --clock-control none --import-source on -k regex:"k_pcg_spmv|k_pcg_update|k_pcg_direction|k_mg_down|k_mg_up|k_mg_resid|k_mg_tail" -s 150 -c 16 -o gpurun_out/r02_pcg_iteration $CMD > gpurun_out/ncu_f.log 2>&1
echo "full rc $?"
tail -2 gpurun_out/ncu_f.log | cut -c1-300
