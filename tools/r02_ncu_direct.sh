#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload dome --steps 3 --warmup 2 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_direct_warp -s 2 -c 1 -o gpurun_out/r02_direct_warp $CMD > gpurun_out/ncu_f.log 2>&1
echo "full rc $?"
