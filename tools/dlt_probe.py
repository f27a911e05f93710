"""Device-resident timing of the DLT kernels on BASELINE config 2 (1M frames x 4 cameras x 10 tracks)."""
import sys; sys.path.insert(0, '/root/repo')
import torch, json
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
st = torch.cuda.Stream(device=dev)
hbm, _ = bench.measured_peaks()
r = bench.bench_dlt(torch, dev, st, 10, 3, 0, 1, hbm)
print({k: r[k] for k in ("points_per_sec", "ms_triangulate", "ms_reproj_error", "ms_fused_reconstruct")})
