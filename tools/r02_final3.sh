#!/bin/bash
# the final code once more: whole GPU suite, smoke, the driver's command on both arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_final3_pytest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r02_final3_pytest.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final3_smoke.log 2>&1; echo "smoke rc $?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final3_bench_grid256.json 2> gpurun_out/r02_final3_bench_grid256.err; echo "bench rc $?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final3_bench_reference.json 2> /dev/null; echo "reference rc $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_final3_bench_grid256.json")); r=json.load(open("gpurun_out/r02_final3_bench_reference.json"))
print("value %.5g e2e %.5g serial %.5g ms %.4f frac %.4f cpu %.4g | reference %.4g" % (d["value"], d["e2e"]["value"], d["e2e"]["serial"], d["ms_per_step"], d["roofline"]["frac"], d["cpu_baseline"]["value"], r["value"]))
for k in d.get("kernels", []): print("   %-50s %8.4f ms %7.0f GB/s share %.3f" % (k["kernel"], k["avg_ms"], k["achieved_gbs"], k["share_of_step"]))
PY
