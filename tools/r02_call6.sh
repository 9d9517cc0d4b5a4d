#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py --no-cpu --steps 10 --no-profile "$@" > gpurun_out/r02_x_$name.json 2> gpurun_out/r02_x_$name.err; echo "$name rc $?"; }
run w3
run w4 --mg-w-levels 4
run w5 --mg-w-levels 5
run w4o15 --mg-w-levels 4 --mg-over-w 1.5
run w4o18 --mg-w-levels 4 --mg-over-w 1.8
run rt1 --pcg-rtol 3e-11 --pcg-rtol-adjoint 1e-7
run rt2 --pcg-rtol 5e-11 --pcg-rtol-adjoint 1e-7
run rt3 --pcg-rtol 3e-11 --pcg-rtol-adjoint 1e-6
run rt1w4 --pcg-rtol 3e-11 --pcg-rtol-adjoint 1e-7 --mg-w-levels 4
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_x_*.json")):
    try:
        d = json.load(open(f))
        print(f, "value %.5g e2e %.5g ms %.3f iters %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"].get("pcg_iterations")), d.get("ms_steps")[:4])
    except Exception as e:
        print(f, "unreadable", e)
PY
python tools/rtol_sweep.py 256 2>&1 | tail -20
