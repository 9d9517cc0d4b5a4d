#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_grid256_n${N}_ring.json 2> gpurun_out/r02_bench_grid256_n${N}_ring.err; echo "rc $?"
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_grid256_n${N}_ring.json"))
print("value %.5g e2e %.5g (serial %.5g) ms %.3f" % (d["value"], d["e2e"]["value"], d["e2e"]["serial"], d["ms_per_step"]), d["config"]["parallelism"], d["config"]["gather_ok"], d["roofline"].get("frac"))
PY
