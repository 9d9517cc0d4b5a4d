#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02_pytest9.log 2>&1; echo "pytest rc $?"
tail -12 gpurun_out/r02_pytest9.log | cut -c1-300
for w in pillow pringle monkey_saddle dome arch; do
  timeout 300 python bench.py --workload $w --no-cpu > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err || echo "$w failed"
  DFDM_DIRECT_WARP=0 timeout 300 python bench.py --workload $w --no-cpu > gpurun_out/r02_bench_${w}_cta.json 2> gpurun_out/r02_bench_${w}_cta.err || echo "$w cta failed"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*.json")):
    if "tail" in f or "n2" in f or "chunks" in f or "_b" in f or "rtol" in f: continue
    try:
        d = json.load(open(f))
        print(f.split("r02_bench_")[1][:-5], "value %.4g e2e %.4g ms %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["roofline"].get("fp64", {}).get("achieved_gflops"))
    except Exception as e:
        print(f, "unreadable", e)
PY
