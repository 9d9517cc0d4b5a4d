#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/two_streams.py 256 2>&1 | tail -5
timeout 300 python tools/two_streams.py 512 2>&1 | tail -5
DFDM_TAIL_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu --no-profile 2>&1 >/dev/null | grep tail-trace | head -16
