#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tail or grid256 or pcg" > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc $?"
tail -5 gpurun_out/r02_pytest5.log
for t in 0 4096; do
  DFDM_MG_TAIL=$t timeout 600 python bench.py --no-cpu --steps 10 > gpurun_out/r02_bench_tail$t.json 2> gpurun_out/r02_bench_tail$t.err; echo "tail $t rc $?"
done
DFDM_TAIL_TRACE=1 python bench.py --steps 1 --warmup 1 --no-cpu --no-profile 2>&1 | grep tail-trace > gpurun_out/r02_trace5.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_tail*.json")):
    try:
        d = json.load(open(f))
        print(f, "value %.5g e2e %.5g ms %.3f iters %s frac %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"].get("pcg_iterations"), d["roofline"]["frac"]), d.get("ms_steps"))
        for k in d.get("kernels", []): print("   %-50s %8.4f ms %7.0f GB/s share %.3f" % (k["kernel"], k["avg_ms"], k["achieved_gbs"], k["share_of_step"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
cat gpurun_out/r02_trace5.txt
