#!/bin/bash
# multigrid cycle shape / over-correction sweep on the default workload (tuning only)
mkdir -p gpurun_out
for wl in -1 1 2 3; do for ov in 1.3 1.5 1.7; do
  python bench.py --steps 2 --warmup 1 --no-cpu --mg-w-levels $wl --mg-over $ov > gpurun_out/cs.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/cs.json")); k={x["kernel"]:x["avg_ms"] for x in d["kernels"]}
print("w_levels $wl over $ov  evals/s %.1f  ms %.1f iters %s coarse %.3f" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations"], k.get("multigrid coarse levels (per V-cycle)",0)))
PY
done; done
