#!/bin/bash
# design pairs per cluster of the small-level kernel (k_mg_bottom): DFDM_BOTTOM_LANES in {8, 16}
for l in ${LANES:-16 8}; do
  DFDM_BOTTOM_LANES=$l timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu 2>/dev/null | LANES_NOW=$l python -c '
import sys, json, os
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("lanes", os.environ["LANES_NOW"], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["config"]["pcg_iterations"], [round(k["avg_ms"], 3) for k in d["kernels"]])
'
done
