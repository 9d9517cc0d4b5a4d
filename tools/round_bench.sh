#!/bin/bash
# the bench lines kept under profiles/: default workload, its two ablations, and the small workloads
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_grid256.json 2> gpurun_out/bench_grid256.err || echo "grid256 failed"
python bench.py --mg-w-levels -1 --no-cpu > gpurun_out/bench_grid256_vcycle.json 2>> gpurun_out/bench_grid256.err || echo "vcycle failed"
for w in pillow pringle monkey_saddle dome arch; do
  python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || echo "$w failed"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("bench_")[1][:-5], d.get("n_gpus"), "value %.4g e2e %.4g ms %.3f cpu %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], (d.get("cpu_baseline") or {}).get("value")))
    except Exception as e:
        print(f, "unreadable", e)
PY
