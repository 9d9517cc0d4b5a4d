#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "direct or golden or batch or vjp or singular or reproducible or arch or tiny" > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc $?"
tail -4 gpurun_out/r02_pytest10.log | cut -c1-300
for w in pillow pringle monkey_saddle dome arch; do
  timeout 300 python bench.py --workload $w --no-cpu > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err || echo "$w failed"
done
python - <<'PY'
import json
for w in ("pillow","pringle","monkey_saddle","dome","arch"):
    try:
        d=json.load(open("gpurun_out/r02_bench_%s.json"%w))
        print(w, "value %.4g e2e %.4g ms %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["roofline"].get("fp64", {}).get("achieved_gflops"))
    except Exception as e: print(w, "unreadable", e)
PY
