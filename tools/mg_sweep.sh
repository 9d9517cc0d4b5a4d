#!/bin/bash
# multigrid parameter sweep on the default workload (tuning only): smoother damping x over-correction of the single-visit levels
mkdir -p gpurun_out
for om in ${OMEGA:-0.7 0.8 0.9}; do for ov in ${OVER:-1.2 1.3 1.4}; do
  python bench.py --steps 1 --warmup 1 --no-cpu --mg-omega $om --mg-over $ov $EXTRA > gpurun_out/mgs.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/mgs.json")); print("omega $om over $ov  evals/s %.1f  ms %.1f iters %s" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations"]))
PY
done; done
