#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02_pytest14.log 2>&1; echo "pytest all rc $?"
tail -4 gpurun_out/r02_pytest14.log | cut -c1-300
DFDM_TEST_TOL_SCALE=0.1 timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r02_pytest14_tight.log 2>&1; echo "tight pytest rc $?"
tail -4 gpurun_out/r02_pytest14_tight.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_p32.json 2> gpurun_out/r02_bench_p32.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_p32.json"))
print("value %.5g ms %.3f e2e %.5g iters %s frac %.4f it-frac %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["pcg_iterations"], d["roofline"]["frac"], d["roofline"]["iteration"]["frac"]), d["config"]["e2e_matches_device"])
for k in d.get("kernels", []): print("   %-50s %8.4f ms %7.0f GB/s share %.3f" % (k["kernel"], k["avg_ms"], k["achieved_gbs"], k["share_of_step"]))
PY
python tools/rtol_sweep.py 2>/dev/null | head -3
