#!/bin/bash
# candidate defaults of the adjoint stop test under the multigrid preconditioner against the parity suite run at tenfold-tightened
# tolerances, and the error against SuperLU on the headline grid
mkdir -p gpurun_out
for v in 5e-9 2e-8 5e-8 1e-7 3e-7; do
  DFDM_PCG_RTOL_ADJOINT_MG=$v DFDM_TEST_TOL_SCALE=0.1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -6 | grep -E "passed|failed|FAILED" | cut -c1-220 | sed "s/^/adjoint $v tight: /"
done
python - <<'PY'
import subprocess, sys, json
for v in ("5e-9", "2e-8", "5e-8", "1e-7", "3e-7"):
    out = subprocess.run([sys.executable, "bench.py", "--steps", "5", "--warmup", "3", "--no-cpu", "--no-profile", "--pcg-rtol-adjoint", v], capture_output=True, text=True).stdout
    d = json.loads(out.strip().splitlines()[-1])
    print("adjoint", v, "iters", d["config"]["pcg_iterations"], "evals/s %.1f" % d["value"])
PY
