#!/bin/bash
# team-per-design sweep on the direct workloads, thread test, streamed e2e on the headline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "direct or golden or batch or vjp or singular or reproducible or arch or tiny or threads" > gpurun_out/r02_pytest11.log 2>&1; echo "pytest rc $?"
tail -4 gpurun_out/r02_pytest11.log | cut -c1-300
DFDM_DIRECT_TEAM_BELOW=100000 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "direct or golden or batch or vjp or singular or reproducible or tiny" > gpurun_out/r02_pytest11_team.log 2>&1; echo "pytest team rc $?"
tail -4 gpurun_out/r02_pytest11_team.log | cut -c1-300
for w in pillow pringle monkey_saddle dome; do
  DFDM_DIRECT_TEAM_BELOW=0 timeout 300 python bench.py --workload $w --no-cpu > gpurun_out/r02_team_${w}_warp.json 2> gpurun_out/r02_team_${w}_warp.err || echo "$w warp failed"
  for t in 64 128 256; do
    DFDM_DIRECT_TEAM_BELOW=100000 DFDM_DIRECT_TEAM_THREADS=$t timeout 300 python bench.py --workload $w --no-cpu > gpurun_out/r02_team_${w}_t$t.json 2> gpurun_out/r02_team_${w}_t$t.err || echo "$w $t failed"
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_team_*.json")):
    try:
        d=json.load(open(f)); print(f.split("r02_team_")[1][:-5], "batch", d["config"]["batch_per_gpu"], "value %.4g e2e %.4g serial %.4g ms %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["serial"], d["ms_per_step"]))
    except Exception as e: print(f, "unreadable", e)
PY
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_stream.json 2> gpurun_out/r02_bench_stream.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_stream.json"))
print("value %.5g ms %.3f e2e" % (d["value"], d["ms_per_step"]), d["e2e"], d["roofline"]["frac"], d["config"]["e2e_matches_device"])
PY
