#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_examples.py tests/test_gpu_api.py -q -m gpu -s > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc $?"
grep "^example \|passed\|failed\|^FAILED\|^E  " gpurun_out/r02_pytest8.log | cut -c1-400 | head -60
