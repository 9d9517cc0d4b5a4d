#!/bin/bash
# round 2, first GPU call: parity suite, stop-test calibration, host-call pipelining
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest.log
tail -5 gpurun_out/r02_pytest.log
timeout 600 python tools/rtol_sweep.py 256 > gpurun_out/r02_rtol_sweep.jsonl 2> gpurun_out/r02_rtol_sweep.err; echo "sweep rc $?"
for c in 1 2 4; do
  DFDM_HOST_CHUNKS=$c timeout 600 python bench.py --no-cpu --no-profile > gpurun_out/r02_bench_chunks$c.json 2> gpurun_out/r02_bench_chunks$c.err; echo "chunks $c rc $?"
done
for b in 64 128; do
  DFDM_HOST_CHUNKS=1 timeout 600 python bench.py --no-cpu --no-profile --batch $b > gpurun_out/r02_bench_b$b.json 2> gpurun_out/r02_bench_b$b.err; echo "batch $b rc $?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, "value %.5g e2e %.5g ms %.3f iters %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"].get("pcg_iterations")))
    except Exception as e:
        print(f, "unreadable", e)
PY
cat gpurun_out/r02_rtol_sweep.jsonl
