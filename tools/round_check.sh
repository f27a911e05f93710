#!/bin/bash
# One GPU call: parity tests, the default bench line, the reference arm, the ncu launch list and a full capture of the
# three largest kernels of the profiled command (tools/profile_target.py, run plainly first).
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_gputests.log; tail -3 gpurun_out/${tag}_gputests.log
python bench.py > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; echo "bench rc=$?"; tail -2 gpurun_out/${tag}_bench_c4.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>/dev/null
python tools/profile_target.py > gpurun_out/${tag}_target.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_target.py > gpurun_out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_schur_mma|k_lin_tile$|k_lin_tile<|k_backsub' -s 3 -c 4 -o gpurun_out/${tag}_full -f python tools/profile_target.py > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
