#!/bin/bash
# over-correction by class of level: --mg-over-w (levels whose coarse problem is solved twice) x --mg-over (the others)
for ot in ${TWICE:-1.3 1.5 1.7}; do for os in ${SINGLE:-1.0 1.15 1.3 1.4}; do
  timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu --mg-over-w $ot --mg-over $os $EXTRA 2>/dev/null | OT=$ot OS=$os python -c '
import sys, json, os
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("twice", os.environ["OT"], "single", os.environ["OS"], round(d["value"], 1), d["config"]["pcg_iterations"], d["config"]["status_ok"])
'
done; done
