#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02_pytest20.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r02_pytest20.log | cut -c1-200
for v in 1 0; do export DFDM_UPDATE_V2=$v;
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_sp2_$v.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r02_sp2_$v.json"))
print("update_v2 $v: %.1f evals/s %s |" % (d["value"], d["config"]["pcg_iterations"]), " ".join("%s %.4f" % (k["kernel"].replace("k_pcg_","").replace("k_mg_","").replace("[level 0]","0").replace("multigrid coarse levels (per V-cycle)","coarse").replace("state, loss and reverse edge kernels (5 launches)","state"), k["avg_ms"]) for k in d["kernels"]))
PY
done
