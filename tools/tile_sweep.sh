#!/bin/bash
mkdir -p gpurun_out
run() {
  DFDM_TILE_SPMV=$1 DFDM_TILE_DOWN=$2 DFDM_TILE_UP=$3 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ts.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/ts.json"))
print("tiles spmv $1 down $2 up $3: %.1f evals/s %s |" % (d["value"], d["config"]["pcg_iterations"]), " ".join("%s %.3f" % (k["kernel"].replace("k_pcg_","").replace("k_mg_","").replace("[level 0]","0").replace("multigrid coarse levels (per V-cycle)","coarse"), k["avg_ms"]) for k in d["kernels"]))
PY
}
run 1 1 1
run 2 1 2
run 4 2 4
run 8 2 8
run 8 4 8
run 16 4 16
run 4 1 4
