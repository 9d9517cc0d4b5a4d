#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02_pytest13.log 2>&1; echo "pytest all rc $?"
tail -4 gpurun_out/r02_pytest13.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_ring.json 2> gpurun_out/r02_bench_ring.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_ring.json"))
print("value %.5g ms %.3f e2e" % (d["value"], d["ms_per_step"]), d["e2e"], d["roofline"]["frac"], d["config"]["e2e_matches_device"], d["cpu_baseline"])
PY
