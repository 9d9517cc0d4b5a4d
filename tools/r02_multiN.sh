#!/bin/bash
# N GPUs: the 2-rank pytest (N >= 2), the driver's bench command, the all-gather variant
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02_final_multi_pytest_n$N.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r02_final_multi_pytest_n$N.log
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc $?"; }
run r02_final_bench_grid256_n${N} bench.py --gpus $N --steps 20 --warmup 5
run r02_final_bench_grid256_n${N}_allgather bench.py --gpus $N --steps 10 --warmup 3 --no-profile --exchange allgather
run r02_final_bench_grid256_n${N}_strong bench.py --gpus $N --steps 10 --warmup 3 --no-profile --scaling strong
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_final_bench_grid256_n${N}*.json")):
    try:
        d = json.load(open(f))
        print(f, "value %.5g e2e %.5g (serial %.5g) ms %.3f" % (d["value"], d["e2e"]["value"], d["e2e"]["serial"], d["ms_per_step"]), d["config"]["parallelism"], d["config"]["gather_ok"], d["scaling"], d["roofline"].get("frac"))
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
