python bench.py --steps 3 --warmup 3 --no-dlt --no-blockdiag --no-e2e --no-extras > gpurun_out/r2_bq.json 2> gpurun_out/r2_bq.err; tail -3 gpurun_out/r2_bq.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bq.json')); print(d['value'], d['ms_per_step'], d['graph_replay']['value']); print(d['roofline']['kernels_ms_per_iteration']); print(d['roofline']['fp64_peaks_tflops'])"
