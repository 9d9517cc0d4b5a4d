#!/bin/bash
# two captures the final set lacked: the level-0 up pass of the multigrid cycle, and the team kernel of the direct path (dome)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --no-cpu --no-profile"
ncu --set full --clock-control none --import-source on -k regex:k_mg_up -s 13 -c 1 -o gpurun_out/r02_mg_up0_final $CMD > gpurun_out/ncu_u.log 2>&1; echo "up0 rc $?"
CMD2="python bench.py --workload dome --steps 3 --warmup 2 --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:k_direct_team -s 2 -c 1 -o gpurun_out/r02_direct_team $CMD2 > gpurun_out/ncu_t.log 2>&1; echo "team rc $?"
