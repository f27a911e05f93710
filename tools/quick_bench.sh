#!/bin/bash
# quick device-timed bench of one workload: bash tools/quick_bench.sh [workload]
wl=${1:-c4}
python bench.py --workload $wl --steps 3 --warmup 3 --no-dlt --no-blockdiag --no-e2e --no-extras > gpurun_out/r2_bq_$wl.json 2> gpurun_out/r2_bq_$wl.err; tail -3 gpurun_out/r2_bq_$wl.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bq_$wl.json')); print('$wl', d['value'], d['ms_per_step'], d['graph_replay']['value']); print(d['roofline']['kernels_ms_per_iteration']); print(d['roofline']['frac'], d['roofline']['iteration'])"
