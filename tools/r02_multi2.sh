#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02_multi2_pytest.log 2>&1; echo "pytest rc $?"; tail -15 gpurun_out/r02_multi2_pytest.log
for ex in allgather root losses mean; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-profile --exchange $ex > gpurun_out/r02_bench_n2_$ex.json 2> gpurun_out/r02_bench_n2_$ex.err; echo "n2 $ex rc $?"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-profile --scaling strong > gpurun_out/r02_bench_n2_strong.json 2> gpurun_out/r02_bench_n2_strong.err; echo "n2 strong rc $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 examples/ensemble_sharded.py > gpurun_out/r02_ensemble_n2.log 2>&1; echo "ensemble rc $?"; tail -2 gpurun_out/r02_ensemble_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/bench_optimizer.py > gpurun_out/r02_optimizer_n2.json 2> gpurun_out/r02_optimizer_n2.err; echo "optimizer rc $?"; cat gpurun_out/r02_optimizer_n2.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_n2_*.json")):
    try:
        d = json.load(open(f))
        print(f, "value %.5g e2e %.5g ms %.3f iters %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"].get("pcg_iterations")), d["config"]["parallelism"], d["config"]["gather_ok"], d["scaling"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
