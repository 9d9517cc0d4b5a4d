#!/bin/bash
# the last line of round 2 on one GPU (both arms) and the ncu capture of the two-design SpMV
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final7_bench_grid256.json 2> gpurun_out/r02_final7_bench_grid256.err; echo "bench rc $?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final7_bench_reference.json 2>/dev/null; echo "reference rc $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_final7_bench_grid256.json")); r=json.load(open("gpurun_out/r02_final7_bench_reference.json"))
print("value %.5g e2e %.5g serial %.5g ms %.4f frac %.4f it %.4f cpu %.4g ref %.4g" % (d["value"], d["e2e"]["value"], d["e2e"]["serial"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["iteration"]["frac"], d["cpu_baseline"]["value"], r["value"]))
for k in d.get("kernels", []): print("   %-50s %8.4f ms %7.0f GB/s share %.3f" % (k["kernel"], k["avg_ms"], k["achieved_gbs"], k["share_of_step"]))
PY
