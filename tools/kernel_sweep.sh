#!/bin/bash
# run a short, iteration-capped bench for every library variant under build/variants (kernel tuning only)
mkdir -p gpurun_out
for so in build/variants/*.so; do
  name=$(basename $so .so)
  DFDM_B200_LIB=$PWD/$so python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sweep_$name.json"))
    print("$name", " ".join("%s %.3f ms %.0f GB/s" % (k["kernel"][6:], k["avg_ms"], k["achieved_gbs"]) for k in d["kernels"]))
except Exception as e:
    print("$name failed", e)
PY
done
