#!/bin/bash
# row-group tiling of the three matrix kernels and the matrix-free product, re-swept for the single-precision p / r copies
mkdir -p gpurun_out
run() {
  DFDM_TILE_SPMV=$1 DFDM_TILE_DOWN=$2 DFDM_TILE_UP=$3 python bench.py --steps 5 --warmup 3 --no-cpu $4 > gpurun_out/ts.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/ts.json"))
print("tiles spmv $1 down $2 up $3 $4: %.1f evals/s %s |" % (d["value"], d["config"]["pcg_iterations"]), " ".join("%s %.4f" % (k["kernel"].replace("k_pcg_","").replace("k_mg_","").replace("[level 0]","0").replace("multigrid coarse levels (per V-cycle)","coarse").replace("state, loss and reverse edge kernels (5 launches)","state"), k["avg_ms"]) for k in d["kernels"]))
PY
}
run 4 1 4
run 4 1 4 --spmv-matrix-free
run 2 1 4
run 8 1 4
run 4 2 4
run 4 4 4
run 4 1 2
run 4 1 8
run 4 1 1
run 8 2 8 --spmv-matrix-free
