#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_last_pytest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r02_last_pytest.log | cut -c1-160
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_last_smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/r02_last_smoke.log
timeout 300 python bench.py --no-cpu > gpurun_out/r02_last_bench_default.json 2> gpurun_out/r02_last_bench_default.err; echo "default bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r02_last_bench_default.json')); print(d['steps'], d['warmup'], round(d['value']), round(d['e2e']['value']), d['roofline']['frac'])"
