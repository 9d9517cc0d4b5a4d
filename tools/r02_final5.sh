#!/bin/bash
# the last lines of round 2: N GPUs (first argument), the driver's command on both arms
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final5_bench_grid256.json 2> gpurun_out/r02_final5_bench_grid256.err; echo "bench rc $?"
  timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final5_bench_reference.json 2>/dev/null; echo "reference rc $?"
  F=gpurun_out/r02_final5_bench_grid256.json
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_final5_bench_grid256_n$N.json 2> gpurun_out/r02_final5_bench_grid256_n$N.err; echo "bench rc $?"
  F=gpurun_out/r02_final5_bench_grid256_n$N.json
fi
python - <<PY
import json
d=json.load(open("$F"))
print("N=%d value %.5g e2e %.5g serial %.5g ms %.4f frac %.4f it %.4f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["e2e"]["serial"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["iteration"]["frac"]), d["config"]["parallelism"], d["config"]["gather_ok"])
for k in d.get("kernels", []): print("   %-50s %8.4f ms %7.0f GB/s share %.3f" % (k["kernel"], k["avg_ms"], k["achieved_gbs"], k["share_of_step"]))
PY
