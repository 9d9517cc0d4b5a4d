#!/bin/bash
# multigrid parameter sweep on the default workload (tuning only)
mkdir -p gpurun_out
for om in 0.7 0.8 0.9; do for ov in 1.3 1.5 1.7 1.9; do
  python bench.py --steps 1 --warmup 1 --no-cpu --mg-omega $om --mg-over $ov --pcg-check-every 1 > gpurun_out/mgs.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/mgs.json")); print("omega $om over $ov  evals/s %.1f  ms %.1f iters %s" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations"]))
PY
done; done
