#!/bin/bash
# after the last host-path changes: the GPU suite, the direct workloads' lines, one headline line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_final2_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02_final2_pytest.log | cut -c1-200
for w in pillow pringle monkey_saddle dome arch; do
  timeout 300 python bench.py --workload $w > gpurun_out/r02_final_bench_$w.json 2> gpurun_out/r02_final_bench_$w.err || echo "$w failed"
done
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_final2_bench_grid256.json 2> gpurun_out/r02_final2_bench_grid256.err; echo "bench rc $?"
python - <<'PY'
import json
for w in ("pillow","pringle","monkey_saddle","dome","arch"):
    d=json.load(open("gpurun_out/r02_final_bench_%s.json"%w))
    print(w, "value %.5g e2e %.5g serial %.5g ms %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["serial"], d["ms_per_step"]), d["roofline"]["kernel"], (d.get("cpu_baseline") or {}).get("value"))
d=json.load(open("gpurun_out/r02_final2_bench_grid256.json"))
print("grid256 value %.5g e2e %.5g serial %.5g ms %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["serial"], d["ms_per_step"]))
PY
