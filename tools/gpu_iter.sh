#!/bin/bash
# development loop on the GPU box: parity tests, device-timed quick benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests.log 2>&1; tail -3 gpurun_out/r2_tests.log
python tools/c1_quick.py 2>&1 | tail -4
for wl in ${@:-c4r8 c3 c4}; do timeout 300 bash tools/quick_bench.sh $wl; done
