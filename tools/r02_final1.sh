#!/bin/bash
# final single-GPU evidence of round 2: tests, smoke, the driver's bench command on both arms, the direct workloads, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02_final_pytest.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench_grid256.json 2> gpurun_out/r02_final_bench_grid256.err; echo "bench rc $?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err; echo "reference rc $?"
for w in pillow pringle monkey_saddle dome arch; do
  timeout 300 python bench.py --workload $w > gpurun_out/r02_final_bench_$w.json 2> gpurun_out/r02_final_bench_$w.err || echo "$w failed"
done
python - <<'PY'
import json
for w in ("grid256","reference","pillow","pringle","monkey_saddle","dome","arch"):
    try:
        d=json.load(open("gpurun_out/r02_final_bench_%s.json"%w))
        print(w, "value %.5g e2e %.5g ms %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("roofline",{}).get("kernel"), d.get("roofline",{}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(w, "unreadable", e)
PY
CMD="python bench.py --steps 1 --warmup 0 --no-cpu --no-profile"
ncu --metrics gpu__time_duration.sum --clock-control none -s 2700 -c 1000 --csv --log-file gpurun_out/r02_grid256_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "launch list rc $?"
ncu --set full --clock-control none --import-source on -k regex:"k_pcg_spmv|k_pcg_update|k_pcg_direction|k_mg_down|k_mg_up|k_mg_tail" -s 144 -c 8 -o gpurun_out/r02_pcg_iteration_final $CMD > gpurun_out/ncu_f.log 2>&1; echo "full rc $?"
