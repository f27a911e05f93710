#!/bin/bash
# development loop on the GPU box: selected tests, device-timed quick benches, one full ncu capture of the Schur kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_vs_reference.py tests/test_gpu_ba.py -m gpu -q > gpurun_out/r2_tests.log 2>&1; tail -3 gpurun_out/r2_tests.log
for wl in c4 c5 c4r8; do bash tools/quick_bench.sh $wl; done
ncu --set full --clock-control none --import-source on -k regex:'k_schur_mma' -s 1 -c 1 -o gpurun_out/r2_schur_full -f python tools/profile_target.py > gpurun_out/r2_ncu_full.log 2>&1
