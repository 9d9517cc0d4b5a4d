#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest7.log 2>&1; echo "pytest rc $?"
tail -15 gpurun_out/r02_pytest7.log
DFDM_TEST_TOL_SCALE=0.1 timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r02_pytest7_tight.log 2>&1; echo "tight pytest rc $?"
tail -15 gpurun_out/r02_pytest7_tight.log
timeout 600 python bench.py --steps 10 > gpurun_out/r02_bench_rtol.json 2> gpurun_out/r02_bench_rtol.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_rtol.json"))
print("value %.5g e2e %.5g ms %.3f iters %s frac %s cpu %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"].get("pcg_iterations"), d["roofline"]["frac"], d.get("cpu_baseline",{}).get("value")), d.get("ms_steps"))
for k in d.get("kernels", []): print("   %-50s %8.4f ms %7.0f GB/s share %.3f" % (k["kernel"], k["avg_ms"], k["achieved_gbs"], k["share_of_step"]))
PY
