#!/bin/bash
# which timed steps are slow, and does the clock sampler (an nvidia-smi loop) have a part in it?
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-profile > gpurun_out/r02_outlier_s$i.json 2> gpurun_out/r02_outlier_s$i.err
  DFDM_BENCH_SAMPLER=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-profile > gpurun_out/r02_outlier_n$i.json 2> gpurun_out/r02_outlier_n$i.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-profile > gpurun_out/r02_outlier_cpu.json 2> gpurun_out/r02_outlier_cpu.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_outlier_*.json")):
    try:
        d=json.load(open(f)); print(f[-12:], round(d["value"]), round(d["e2e"]["value"]), d["ms_steps"], [round(x) for x in d["ms_steps_host_enqueue"]])
    except Exception as e: print(f, e)
PY
