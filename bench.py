#!/usr/bin/env python
"""bench.py - throughput of the bundle-adjustment hot path (and the DLT path) on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      (the reference's own CPU implementation, oracle/_ref, on a bounded sample)

Workload (BASELINE.json config 4): 16 cameras, 2,000,000 points x 10 views = 20,000,000 observations, all 10 Argus
intrinsics + pose free (nccalib = ncdist = 0).  With N > 1 the points (and all their observations) of THAT scene are sharded
over the ranks, balanced by observations, cameras replicated - the configuration as BASELINE.json words it ("sharded across
2/4/8 B200"): strong scaling, the default.  `--scaling weak` gives every rank its own 20 M-observation scene instead.
A "step" is one pass of the hot path over that batch: ITERS_PER_STEP (10) Levenberg-Marquardt iterations from the same
perturbed start (stop tests disabled so every step does identical work; asserted).
metric = observations x LM iterations per second, whole job (all ranks).  Per damping attempt the ranks exchange one
all-reduce of [U, ea, Schur sums, E] + two maxima and one of four scalars (NCCL, issued by the library itself).

One JSON line on rank 0; see DESIGN.md for the roofline definitions.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP64_PEAK_FALLBACK = 37.1    # only if the in-run measurement (argus_fp64_peak) fails: profiles/r01_fp64_peak.txt
ITERS_PER_STEP = 10
WORKLOADS = {
    # name: (cameras, points per rank, views per point, nccalib, ncdist)
    "c4": (16, 2_000_000, 10, 0, 0),
    "c3": (8, 200_000, 6, 0, 0),
    "small": (8, 20_000, 6, 0, 0),
    "c4r8": (16, 250_000, 10, 0, 0),       # what one rank of an 8-GPU strong-scaling run of c4 holds (fixed per-iteration costs)
    # config 5: 32 cameras, 5M points x 12 views, 10 % outlier tracks, cameras 0-15 fixed intrinsics / 16-31 free (extension)
    "c5": (32, 5_000_000, 12, 0, 0),
}
OPTS_FIXED_WORK = [1e-3, 0.0, 0.0, 0.0, 0.0]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index; self.proc = None; self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def make_scene(workload, rank):
    from argus_gui_b200 import scenes
    m, n, views, nk, nd = WORKLOADS[workload]
    s = scenes.make_ba_scene(m, n, views, seed=4 + 1000 * rank, cam_seed=4, outlier_frac=0.1 if workload == "c5" else 0.0,
                             radius=9.0 if workload == "c5" else 8.0)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    return s, p0, rot0, (m, n, views, nk, nd)


def measure_fp64_peak(device_index):
    """fp64 throughput of this GPU, measured now (MEASURED_PEAKS.json has no fp64 entry): DMMA m8n8k4 and DFMA, TFLOP/s."""
    import ctypes
    from argus_gui_b200 import _lib
    a, b = ctypes.c_double(0.0), ctypes.c_double(0.0)
    rc = _lib.load().argus_fp64_peak(int(device_index), 0.3, ctypes.byref(a), ctypes.byref(b))
    if rc != 0 or a.value <= 0:
        return FP64_PEAK_FALLBACK, FP64_PEAK_FALLBACK, "fallback constant (profiles/r01_fp64_peak.txt)"
    return a.value, b.value, "measured in this run (argus_fp64_peak: mma.sync m8n8k4 f64 / DFMA loops, 0.3 s each)"


def algorithmic_counts(vmask, m):
    """per LM iteration flops = nobs*(1025 + 90 + 1316) + 768 * sum_i v_i (v_i + 1) + Sdim^3/3,
    bytes = 40 nobs + 72 npts + 8 Sdim^2; the Schur kernel alone: 768 * sum_i v_i (v_i + 1) flops."""
    v = vmask.sum(axis=1).astype(np.float64)
    nobs = float(v.sum()); n = float(vmask.shape[0]); sdim = 16.0 * m
    schur = 768.0 * float((v * (v + 1.0)).sum())
    return {"nobs": nobs, "flops_iter": nobs * 2431.0 + schur + sdim ** 3 / 3.0, "bytes_iter": 40.0 * nobs + 72.0 * n + 8.0 * sdim * sdim,
            "schur_flops": schur}


def reference_sample(s, p0, m, n_sample):
    """First n_sample points of the scene with all their observations (same cameras): the bounded CPU sample."""
    n_sample = min(n_sample, s["vmask"].shape[0])
    vm = s["vmask"][:n_sample]
    nobs = int(vm.sum())
    return vm, s["x"][:nobs], np.concatenate([p0[:16 * m], p0[16 * m:16 * m + 3 * n_sample]])


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own C code (oracle/_ref: sba 1.6 + libsbaprojs + OpenBLAS, single thread - the
    reference is single-threaded and more BLAS threads slow it down) on a bounded sample of the workload."""
    if rank != 0:
        return
    from oracle import sba_ref, ba_port
    s, p0, rot0, (m, n, views, nk, nd) = make_scene(args.workload, 0)
    vm, x, p = reference_sample(s, p0, m, args.ref_points)
    nobs = int(vm.sum())
    if sba_ref.available():
        kind = "reference"
        run = lambda: sba_ref.ref_motstr(p, x, vm, rot0, nk, nd, itmax=ITERS_PER_STEP, opts=OPTS_FIXED_WORK)
    else:
        kind = "port"
        run = lambda: ba_port.port_motstr(p, x, vm, rot0, nk, nd, itmax=ITERS_PER_STEP, opts=OPTS_FIXED_WORK)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, info, ret = run()
        assert int(info[5]) == ITERS_PER_STEP, info
    dt = time.perf_counter() - t0
    value = nobs * ITERS_PER_STEP * args.steps / dt
    sample = "first %d points (%d observations) of the %s scene, %d cameras, %d LM iterations per step" % (vm.shape[0], nobs, args.workload, m, ITERS_PER_STEP)
    line = {"impl": "reference", "metric": "ba_observations_per_second", "value": value, "unit": "obs*iters/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: BA %d cameras, bounded sample" % (args.workload, m), "sample": sample, "iters_per_step": ITERS_PER_STEP},
            "cpu_baseline": {"value": value, "unit": "obs*iters/s", "cores": 1, "kind": kind, "sample": sample, "host_cores": os.cpu_count()},
            "lm_iters_per_sec": ITERS_PER_STEP * args.steps / dt,
            "e2e": {"value": value, "unit": "obs*iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bench_blockdiag(torch, ba, stream, local, s, p0, rot0, m, nk, nd, steps, cpu_points):
    """Motion-only / structure-only LM on the same scene: device-resident throughput by CUDA events and the
    reference's sba_mot_levmar_x / sba_str_levmar_x on a bounded sample of the same scene (rank 0, N = 1)."""
    out = {}
    for mode, iters in (("mot", 10), ("str", 3)):
        with torch.cuda.stream(stream):
            B = ba.BundleAdjuster(device=local, stream=stream.cuda_stream, mode=mode)
            B.upload(s["vmask"], p0, rot0, s["x"], nk, nd)
            for _ in range(2):
                B.reset(); ret, info = B.solve(itmax=iters, opts=OPTS_FIXED_WORK)
            c0 = B.counters()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                B.reset(); ret, info = B.solve(itmax=iters, opts=OPTS_FIXED_WORK)
                assert ret == iters and int(info[5]) == iters, (mode, ret, info)
            e1.record(stream); e1.synchronize()
            ms = e0.elapsed_time(e1) / steps
            c1 = B.counters()
            B.close()
        out[mode] = {"obs_iters_per_sec": s["nobs"] * iters / (ms * 1e-3), "ms_per_iteration": ms / iters, "iters_per_step": iters,
                     "linear_systems": int(info[9]), "gpu_launches": int(c1["launches"] - c0["launches"])}
    if cpu_points:
        from oracle import sba_ref, ba_port
        vm, xs, ps = reference_sample(s, p0, m, cpu_points)
        for mode, iters in (("mot", 10), ("str", 3)):
            t0 = time.perf_counter()
            if sba_ref.available() and mode == "mot":
                _, iref, _ = sba_ref.ref_mot(ps[:16 * m], ps[16 * m:], xs, vm, rot0, nk, nd, itmax=iters, opts=OPTS_FIXED_WORK); kind = "reference"
            elif sba_ref.available():
                # the reference's structure-only callbacks read the full camera quaternion from the camera row (no rot0)
                cf = ps[:16 * m].reshape(m, 16).copy(); q = np.asarray(rot0).reshape(m, 4); cf[:, 10:13] = q[:, 1:4] * np.sign(q[:, :1] + (q[:, :1] == 0))
                _, iref, _ = sba_ref.ref_str(cf.ravel(), ps[16 * m:], xs, vm, itmax=iters, opts=OPTS_FIXED_WORK); kind = "reference"
            else:
                fn = ba_port.port_mot if mode == "mot" else ba_port.port_str
                _, iref, _ = (fn(ps, xs, vm, rot0, nk, nd, itmax=iters, opts=OPTS_FIXED_WORK) if mode == "mot"
                              else fn(ps, xs, vm, rot0, itmax=iters, opts=OPTS_FIXED_WORK)); kind = "port"
            dt = time.perf_counter() - t0
            out[mode]["cpu_baseline"] = {"value": int(vm.sum()) * int(iref[5]) / dt, "unit": "obs*iters/s", "cores": 1, "kind": kind,
                                         "sample": "first %d points of the same scene, %d LM iterations, %.1f s" % (vm.shape[0], int(iref[5]), dt)}
    return out


def bench_pair(rank):
    """Two-view initialisation (Triangulator.triangulate of one camera pair): 1M point pairs through the host
    API (argus_pair_triangulate: H2D, undistortion, 8-point F, 4 hypotheses x N two-view solves, vote, final cloud, D2H), and
    the CPU restatement (vectorised numpy + OpenCV; the reference itself loops in Python) on a bounded sample."""
    from argus_gui_b200 import scenes, triangulate
    n = 1_000_000
    s = scenes.make_ba_scene(2, n, 2, seed=5, radius=5.0)
    dense = np.ascontiguousarray(s["x"].reshape(n, 4))
    intr = s["cams_true"][:, :10]
    mk = lambda mod, sl: mod.Triangulator(dense[sl, :2], dense[sl, 2:], intr[0, 0], intr[1, 0], intr[0, 1:3], intr[1, 1:3], intr[0, -5:], intr[1, -5:])
    mk(triangulate, slice(0, n)).triangulate()
    t0 = time.perf_counter()
    for _ in range(3):
        mk(triangulate, slice(0, n)).triangulate()
    dt = (time.perf_counter() - t0) / 3
    out = {"points_per_sec": n / dt, "ms": 1e3 * dt, "sample": "%d point pairs, one camera pair, host API" % n}
    if rank == 0:
        from oracle import tri_port
        ns = 50_000
        t0 = time.perf_counter(); mk(tri_port, slice(0, ns)).triangulate(); dtc = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": ns / dtc, "unit": "points/s", "cores": 1, "kind": "port",
                               "sample": "first %d point pairs (vectorised numpy + OpenCV restatement of triangulate.py), %.1f s" % (ns, dtc)}
    # per-camera DLT fit + reprojection distances (tools.solve_dlt / getCoefficients), 2M points, host API
    from argus_gui_b200 import dlt as dlt_host
    nf = 2_000_000
    sc = scenes.make_dlt_scene(nf, m=2, ntracks=1, seed=9, nan_frac=0.0)
    xyz = sc["xyz_true"].reshape(nf, 3); L0 = sc["dlt"][0]
    den = xyz @ L0[8:11] + 1.0
    uv = np.stack([(xyz @ L0[0:3] + L0[3]) / den, (xyz @ L0[4:7] + L0[7]) / den], axis=1) + np.random.default_rng(1).normal(0, 0.5, (nf, 2))
    dlt_host.fit(xyz, uv)
    t0 = time.perf_counter()
    for _ in range(3):
        dlt_host.fit(xyz, uv)
    dtf = (time.perf_counter() - t0) / 3
    out["dlt_fit"] = {"points_per_sec": nf / dtf, "ms": 1e3 * dtf, "sample": "%d points, one camera, host API (H2D + QR tree + distances + D2H)" % nf}
    if rank == 0:
        from oracle import dlt_port
        ns = 200_000
        t0 = time.perf_counter(); dlt_port.solve_dlt(xyz[:ns], uv[:ns]); dtc = time.perf_counter() - t0
        out["dlt_fit"]["cpu_baseline"] = {"value": ns / dtc, "unit": "points/s", "cores": os.cpu_count(), "kind": "port",
                                          "sample": "first %d points, numpy lstsq (vectorised; the reference fills the matrix in a Python loop), %.1f s" % (ns, dtc)}
    return out


def bench_dlt(torch, device, stream, steps, warmup, rank, world, hbm_peak, dfma_peak):
    """BASELINE config 2 per rank: 1M frames x 4 cameras x 10 tracks; triangulation + reprojection error."""
    import ctypes
    from argus_gui_b200 import scenes, dlt as dlt_host, _lib
    lib = _lib.load()
    N, m, k = 1_000_000, 4, 10
    d = scenes.make_dlt_scene(N, m=m, ntracks=k, seed=2 + rank)
    with torch.cuda.stream(stream):
        t_pts = torch.from_numpy(d["pts"]).to(device)
        t_prof = torch.from_numpy(np.ascontiguousarray(d["prof"])).to(device)
        t_dlt = torch.from_numpy(np.ascontiguousarray(d["dlt"])).to(device)
        t_xyz = torch.empty((N, 3 * k), dtype=torch.float64, device=device)
        t_err = torch.empty((k, N), dtype=torch.float64, device=device)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)   # > 126 MB L2
        vp = ctypes.c_void_p
        sp = vp(stream.cuda_stream)

        def step():
            rc = lib.argus_dlt_triangulate_dev(vp(t_pts.data_ptr()), N, m, k, vp(t_prof.data_ptr()), vp(t_dlt.data_ptr()), vp(t_xyz.data_ptr()), 0, sp)
            rc |= lib.argus_dlt_reproj_error_dev(vp(t_xyz.data_ptr()), vp(t_pts.data_ptr()), N, m, k, vp(t_prof.data_ptr()), vp(t_dlt.data_ptr()),
                                                 vp(t_err.data_ptr()), sp)
            assert rc == 0
        for _ in range(warmup):
            step()
        tri_ms = err_ms = fused_ms = 0.0
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for _ in range(steps):
            flush.zero_()
            e[0].record(stream)
            lib.argus_dlt_triangulate_dev(vp(t_pts.data_ptr()), N, m, k, vp(t_prof.data_ptr()), vp(t_dlt.data_ptr()), vp(t_xyz.data_ptr()), 0, sp)
            e[1].record(stream)
            lib.argus_dlt_reproj_error_dev(vp(t_xyz.data_ptr()), vp(t_pts.data_ptr()), N, m, k, vp(t_prof.data_ptr()), vp(t_dlt.data_ptr()),
                                           vp(t_err.data_ptr()), sp)
            e[2].record(stream)
            lib.argus_dlt_reconstruct_dev(vp(t_pts.data_ptr()), N, m, k, vp(t_prof.data_ptr()), vp(t_dlt.data_ptr()), vp(t_xyz.data_ptr()),
                                          vp(t_err.data_ptr()), 0, sp)
            e[3].record(stream)
            e[3].synchronize()
            tri_ms += e[0].elapsed_time(e[1]); err_ms += e[1].elapsed_time(e[2]); fused_ms += e[2].elapsed_time(e[3])
    tri_ms /= steps; err_ms /= steps; fused_ms /= steps
    npts = N * k
    tri_bytes = npts * (2 * m * 8 + 24); err_bytes = npts * (2 * m * 8 + 24 + 8)
    # end to end through the host API (pinned numpy in, pinned numpy out; the library pipelines H2D / kernel / D2H by chunks)
    h_pts = torch.from_numpy(d["pts"]).pin_memory().numpy()
    h_xyz = torch.empty((N, 3 * k), dtype=torch.float64).pin_memory().numpy()
    h_err = torch.empty((k, N), dtype=torch.float64).pin_memory().numpy()
    dlt_host.reconstruct(h_pts, d["prof"], d["dlt"], out=(h_xyz, h_err))
    n_e2e = 3
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        xyz, err = dlt_host.reconstruct(h_pts, d["prof"], d["dlt"], out=(h_xyz, h_err))
    e2e_s = (time.perf_counter() - t0) / n_e2e
    fused_bytes = npts * (2 * m * 8 + 24 + 8)
    out = {"workload": "c2: %d frames x %d cameras x %d tracks per rank" % (N, m, k), "points_per_sec": npts / (fused_ms * 1e-3),
           "points_per_sec_two_calls": npts / ((tri_ms + err_ms) * 1e-3),
           "ms_triangulate": tri_ms, "ms_reproj_error": err_ms, "ms_fused_reconstruct": fused_ms,
           # about 900 fp64 operations per (frame, track) with 4 cameras: 4 x 5 undistortion iterations ~ 700, the (n+1)-row least
           # squares ~ 110, the reprojection ~ 90; ncu: fp64 pipe 71-77 % busy, DRAM 13-17 % (profiles/r01_final_ncu_full_dlt.txt)
           "roofline_fused": {"bound": "fp64", "achieved": 900.0 * npts / (fused_ms * 1e-3) / 1e12, "peak": dfma_peak, "unit": "TFLOP/s",
                              "frac": 900.0 * npts / (fused_ms * 1e-3) / 1e12 / dfma_peak, "traffic": None,
                              "hbm_frac": fused_bytes / (fused_ms * 1e-3) / 1e9 / hbm_peak,
                              "note": "fp64-instruction bound: the executed instruction stream keeps the fp64 pipe 71-77 % busy (ncu), about 4x the algorithmic count above - "
                                      "5 undistortion iterations with a correctly rounded division each and no FMA contraction, the price of bit-exact float32 output"},
           "roofline_triangulate": {"bound": "fp64", "achieved": 810.0 * npts / (tri_ms * 1e-3) / 1e12, "peak": dfma_peak, "unit": "TFLOP/s",
                                    "frac": 810.0 * npts / (tri_ms * 1e-3) / 1e12 / dfma_peak, "hbm_frac": tri_bytes / (tri_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None},
           "roofline_reproj_error": {"bound": "fp64", "achieved": 790.0 * npts / (err_ms * 1e-3) / 1e12, "peak": dfma_peak, "unit": "TFLOP/s",
                                     "frac": 790.0 * npts / (err_ms * 1e-3) / 1e12 / dfma_peak, "hbm_frac": err_bytes / (err_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None},
           "e2e": {"value": npts / e2e_s, "unit": "points/s", "h2d_bytes_per_step": int(d["pts"].nbytes), "d2h_bytes_per_step": int(xyz.nbytes + err.nbytes)},
           "gpu_launches": 3 * steps}
    # bootstrap confidence intervals (bootstrapXYZs: 250 perturbed re-triangulations per point), host API, 100k frames
    nb = 100_000
    rm = np.ascontiguousarray(np.nan_to_num(err[:, :nb].T, nan=1.0))
    t0 = time.perf_counter()
    sd = dlt_host.bootstrap_sd(np.ascontiguousarray(d["pts"][:nb]), rm, d["prof"], d["dlt"], bs_iter=250, seed=1)
    bt = time.perf_counter() - t0
    out["bootstrap"] = {"points_per_sec": nb * k / bt, "retriangulations_per_sec": nb * k * 250 / bt, "bs_iter": 250,
                        "sample": "%d frames x %d tracks through argus_dlt_bootstrap (host buffers)" % (nb, k)}
    if rank == 0 and world == 1:
        from oracle import dlt_port
        ns = 1500
        t0 = time.perf_counter()
        for t in range(k):
            xo = dlt_port.uv_to_xyz(d["pts"][:ns, t * 2 * m:(t + 1) * 2 * m], d["prof"], d["dlt"])
        eo = dlt_port.get_repo_errors(xyz[:ns], d["pts"][:ns], d["prof"], d["dlt"])
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": ns * k / dt, "unit": "points/s", "cores": 1, "kind": "port",
                               "sample": "first %d frames x %d tracks (numpy restatement of tools.uv_to_xyz + get_repo_errors)" % (ns, k),
                               "note": "the reference's tools.py does not travel to the GPU box; timed in the build container on the same sample "
                                       "(tools/dlt_reference_timing.py) it does 3.6e3 points/s against this restatement's 1.7e4, with bit-identical output"}
    del t_pts, t_xyz, t_err, flush
    return out


def bench_extras(torch, ba, local, stream, args, dmma_peak):
    """BASELINE.json's other named configurations and the reference's own known-answer run, N = 1, rank 0 (short runs beside
    the headline): c1 through the wand driver with the reference's C code timed on the same sba inputs, c3, c5 (device timed,
    10 LM iterations per step) and sba-1.6p/demo's 54camsvarKD data set (published: 34 iterations, error 0.128779, 55.17 s)."""
    import contextlib
    import io
    import tempfile
    from argus_gui_b200 import scenes, sbaDriver, sba
    from oracle import sba_ref
    out = {}
    # ---- c1: 3-camera wand calibration through sbaArgusDriver.fix() (what Argus users run) ----
    w = scenes.make_wand_scene(nframes=2000, nbackground=100, m=3, seed=1)
    tmp = tempfile.mkdtemp()
    sink = io.StringIO()
    runs = []
    for rep in range(3):
        with contextlib.redirect_stdout(sink):
            t0 = time.perf_counter()
            d = sbaDriver.sbaArgusDriver(w["ppts"].copy(), w["uppts"].copy(), w["cams"].copy(), display=False, scale=0.5, modeString="01",
                                         name=os.path.join(tmp, "c1"), temp=tmp, report=False)
            t1 = time.perf_counter()
            res = d.fix()
            t2 = time.perf_counter()
        runs.append((t1 - t0, t2 - t1))
    points = sba.Points.fromDylan(d.pts); cameras = sba.Cameras.fromDylan(np.hstack((d.cams, d.ext)))
    a, r0 = cameras.pack()
    p_sba = np.concatenate([a.ravel(), points.B.ravel()])
    nobs1 = int(points.vmask.sum())
    with torch.cuda.stream(stream):
        B = ba.BundleAdjuster(device=local, stream=stream.cuda_stream)
        B.upload(points.vmask, p_sba, r0.ravel(), points.X, 5, 4)
        for _ in range(3):
            B.reset(); B.solve(itmax=150)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10):
            B.reset(); ret1, info1 = B.solve(itmax=150)
        torch.cuda.synchronize(); res_ms = 1e3 * (time.perf_counter() - t0) / 10
        B.close()
    t0 = time.perf_counter(); ba.solve_motstr(points.vmask, p_sba, r0.ravel(), points.X, 5, 4, itmax=150); one_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter(); ba.solve_motstr(points.vmask, p_sba, r0.ravel(), points.X, 5, 4, itmax=150); one_ms = min(one_ms, 1e3 * (time.perf_counter() - t0))
    c1 = {"workload": "c1: 3 cameras, 2000 wand frames + 100 background points -> %d points, %d observations, mode '01'" % (d.pts.shape[0], nobs1),
          "driver_ms": {"constructor (parse + CUDA two-view initialisation)": 1e3 * min(r[0] for r in runs), "fix() (BA + DLT fits + files)": 1e3 * min(r[1] for r in runs)},
          "ba_iterations": int(ret1), "ba_resident_ms": res_ms, "ba_ms_per_iteration": res_ms / max(int(ret1), 1), "ba_one_shot_call_ms": one_ms,
          "wand_score": float(res.wandscore), "dlt_rmse": [float(v) for v in res.dlterrors]}
    if sba_ref.available():
        t0 = time.perf_counter(); pr, ir, rr = sba_ref.ref_motstr(p_sba, points.X, points.vmask, r0.ravel(), 5, 4, itmax=150); dt = time.perf_counter() - t0
        c1["cpu_baseline"] = {"value": 1e3 * dt, "unit": "ms for the same bundle adjustment (%d iterations)" % rr, "cores": 1, "kind": "reference",
                              "sample": "the whole problem (sba_motstr_levmar_x + img_projsKDRTS_x/_jac_x on the same arrays)"}
        c1["final_cost_rel_diff_vs_reference"] = abs(info1[1] - ir[1]) / ir[1]
    out["c1"] = c1

    # ---- c3 / c5: device-timed, ITERS_PER_STEP iterations per step ----
    for wl in (["c3"] + ([] if args.no_c5 else ["c5"])):
        s, p0, rot0, (m, n, views, nk, nd) = make_scene(wl, 0)
        cnt = algorithmic_counts(s["vmask"], m)
        with torch.cuda.stream(stream):
            B = ba.BundleAdjuster(device=local, stream=stream.cuda_stream)
            B.upload(s["vmask"], p0, rot0, s["x"], nk, nd)
            if wl == "c5":
                B.set_camera_fix([5] * 16 + [0] * 16, [5] * 16 + [0] * 16)
            for _ in range(2):
                B.reset(); B.solve(itmax=ITERS_PER_STEP, opts=OPTS_FIXED_WORK)
            steps = 3
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            prof = {}
            e0.record(stream)
            for _ in range(steps):
                B.reset(); ret, info = B.solve(itmax=ITERS_PER_STEP, opts=OPTS_FIXED_WORK, profile=True)
                assert ret == ITERS_PER_STEP, (wl, ret, info)
                for kname, (ms, c_) in B.profile().items():
                    q = prof.setdefault(kname, [0.0, 0]); q[0] += ms; q[1] += c_
            e1.record(stream); e1.synchronize()
            ms_it = e0.elapsed_time(e1) / (steps * ITERS_PER_STEP)
            B.close()
        sch = prof["schur"][0] / max(prof["schur"][1], 1)
        r = {"workload": "%s: %d cameras, %d points x %d views = %d observations%s" % (wl, m, n, views, s["nobs"],
                         ", cameras 0-15 fixed intrinsics / 16-31 free, 10 % outlier tracks" if wl == "c5" else ""),
             "value": s["nobs"] / (ms_it * 1e-3), "unit": "obs*iters/s", "ms_per_iteration": ms_it,
             "kernels_ms_per_iteration": {k_: v[0] / v[1] for k_, v in sorted(prof.items()) if v[1] and k_ not in ("residual", "linearize_stats")},
             "schur": {"ms_per_launch": sch, "tflops": cnt["schur_flops"] / (sch * 1e-3) / 1e12, "frac_of_fp64_peak": cnt["schur_flops"] / (sch * 1e-3) / 1e12 / dmma_peak},
             "iteration_fp64_frac": cnt["flops_iter"] / (ms_it * 1e-3) / 1e12 / dmma_peak}
        if sba_ref.available() and wl == "c3":
            vm, xs, ps = reference_sample(s, p0, m, 16384)
            t0 = time.perf_counter(); _, iref, _ = sba_ref.ref_motstr(ps, xs, vm, rot0, nk, nd, itmax=ITERS_PER_STEP, opts=OPTS_FIXED_WORK); dt = time.perf_counter() - t0
            r["cpu_baseline"] = {"value": int(vm.sum()) * int(iref[5]) / dt, "unit": "obs*iters/s", "cores": 1, "kind": "reference",
                                 "sample": "first %d points (%d observations) of the same scene, %d LM iterations, %.1f s" % (vm.shape[0], int(vm.sum()), int(iref[5]), dt)}
        if wl == "c5":
            r["cpu_baseline"] = {"value": None, "kind": "reference", "sample": "not runnable: the reference indexes its Jacobian buffer with a 32-bit int (sba_levmar.c:676), "
                                 "60 M observations overflow it, and it has no per-camera fixed / free masks"}
        out[wl] = r
        del s, p0

    # ---- c3 as BASELINE.json words it: "... + wand-length constraint" - the constraint INSIDE the solver (extension, parity
    #      unpinned: the reference only rescales afterwards).  Same shape: 8 cameras, 100 000 wand positions = 200 000 points ----
    m, n, views, nk, nd = WORKLOADS["c3"]
    sw = scenes.make_ba_wand_scene(m, n // 2, 0, views, seed=4, cam_seed=4)
    p0, rot0 = scenes.pack_ba_problem(sw["cams0"], sw["pts0"])
    with torch.cuda.stream(stream):
        B = ba.BundleAdjuster(device=local, stream=stream.cuda_stream)
        B.upload(sw["vmask"], p0, rot0, sw["x"], nk, nd)
        res_w = {}
        for label, weight in (("off", 0.0), ("on", 1000.0)):
            B.set_wand(sw["ia"], sw["ib"], sw["wand"], weight)
            for _ in range(2):
                B.reset(); B.solve(itmax=ITERS_PER_STEP, opts=OPTS_FIXED_WORK)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                B.reset(); ret, info = B.solve(itmax=ITERS_PER_STEP, opts=OPTS_FIXED_WORK)
                assert ret == ITERS_PER_STEP, (label, ret, info)
            e1.record(stream); e1.synchronize()
            ms_it = e0.elapsed_time(e1) / (3 * ITERS_PER_STEP)
            B.reset(); retc, infoc = B.solve(itmax=150)
            pts = B.download()[16 * m:].reshape(-1, 3)
            d = np.linalg.norm(pts[sw["ia"]] - pts[sw["ib"]], axis=1)
            res_w[label] = {"ms_per_iteration": ms_it, "obs_iters_per_sec": sw["nobs"] / (ms_it * 1e-3), "iterations_to_convergence": int(retc),
                            "stop_reason": int(infoc[6]), "final_cost": float(infoc[1]),
                            "wand_score_percent": float(100.0 * d.std() / d.mean()), "mean_wand_length": float(d.mean())}
        B.close()
    out["c3_wand"] = {"workload": "c3 with the wand-length constraint in the solver: %d cameras, %d wand positions = %d points x %d views = %d observations, "
                                  "weight 1000 px per length unit, true length %.2f" % (m, n // 2, n, views, sw["nobs"], sw["wand"]),
                      "parity": "unpinned (extension: the reference applies the wand length after the adjustment, sbaDriver.py:703-713)",
                      "constraint_off": res_w["off"], "constraint_on": res_w["on"]}
    del sw, p0

    # ---- the reference's own largest known-answer run: demo/54camsvarKD.txt + 54pts.txt (README.txt:197-206) ----
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    if helpers.have_demo():
        cams, pts, vmask, x = helpers.load_demo("54camsvarKD.txt", "54pts.txt")
        p0, rot0 = scenes.pack_ba_problem(cams, pts)
        ba.solve_motstr(vmask, p0, rot0, x, 2, 3, itmax=150)
        t0 = time.perf_counter(); pg, ig, rg = ba.solve_motstr(vmask, p0, rot0, x, 2, 3, itmax=150); dtg = time.perf_counter() - t0
        k = {"workload": "sba-1.6p/demo 54camsvarKD: 54 cameras, 5207 points, 24609 projections, 16485 variables",
             "published": {"iterations": 34, "final_error": 0.128779, "elapsed_s": 55.17, "source": "sba-1.6p/demo/README.txt:197-206 (hardware not stated)"},
             "gpu": {"iterations": int(rg), "stop_reason": int(ig[6]), "final_error": float(ig[1] / x.shape[0]), "one_shot_call_s": dtg}}
        if sba_ref.available():
            t0 = time.perf_counter(); pr, ir, rr = sba_ref.ref_motstr(p0, x, vmask, rot0, 2, 3, itmax=150); dtr = time.perf_counter() - t0
            k["cpu_baseline"] = {"value": dtr, "unit": "s", "cores": 1, "kind": "reference", "iterations": int(rr), "final_error": float(ir[1] / x.shape[0]),
                                 "sample": "the whole problem, the reference's own compiled code on this host"}
        out["demo_54camsvarKD"] = k
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the ONE scene of the workload sharded over the ranks (BASELINE.json config 4 as written); weak = one scene per rank")
    ap.add_argument("--ref-points", type=int, default=4096, help="points in the bounded CPU sample")
    ap.add_argument("--cpu-baseline-points", type=int, default=32768)
    ap.add_argument("--no-dlt", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-blockdiag", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the c1 / c3 / c5 / 54camsvarKD section (N = 1)")
    ap.add_argument("--no-c5", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from argus_gui_b200 import ba
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the BA / DLT path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    hbm_peak, hbm_src = measured_peaks()
    dmma_peak, dfma_peak, fp64_src = measure_fp64_peak(local)
    stream = torch.cuda.Stream(device=device)

    strong = args.scaling == "strong"
    s, p0, rot0, (m, n, views, nk, nd) = make_scene(args.workload, 0 if strong else rank)
    counts = algorithmic_counts(s["vmask"], m)              # of the whole scene (strong) / of this rank's scene (weak)
    if strong and world > 1:
        vm_l, p_l, x_l, (lo, hi) = ba.shard_problem(s["vmask"], p0, s["x"], m, world, rank, balance_observations=True)
        vm_l = np.ascontiguousarray(vm_l); p_l = np.ascontiguousarray(p_l); x_l = np.ascontiguousarray(x_l)
    else:
        vm_l, p_l, x_l = s["vmask"], p0, s["x"]
    n_l = vm_l.shape[0]
    nobs_local = int(np.count_nonzero(vm_l))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def new_adjuster():
        B_ = ba.BundleAdjuster(device=local, stream=stream.cuda_stream)
        if world > 1:
            B_.init_nccl(rank, world)
        return B_

    with torch.cuda.stream(stream):
        B = new_adjuster()
        B.upload(vm_l, p_l, rot0, x_l, nk, nd)
        if args.workload == "c5":
            B.set_camera_fix([5] * 16 + [0] * 16, [5] * 16 + [0] * 16)
        for _ in range(args.warmup):
            B.reset(); ret, info = B.solve(itmax=ITERS_PER_STEP, opts=OPTS_FIXED_WORK)
            assert ret == ITERS_PER_STEP, (ret, info)
        c0 = B.counters()
        sampler = ClockSampler(local); sampler.start()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        # N = 1: the per-phase CUDA events are taken INSIDE the timed region (event-record nodes in every fourth attempt's graph;
        # the cost is below 0.1 %).  N > 1: the host wait in front of such an attempt lets the ranks drift apart before a
        # collective (measured: 7 % at 2 GPUs), so the timed region is the product path as it is and the per-phase events come
        # from a second pass of the same steps right after it.
        prof = {}
        prof_in_timed = world == 1

        def run_steps(profile):
            last = None
            for _ in range(args.steps):
                B.reset(); ret_, info_ = B.solve(itmax=ITERS_PER_STEP, opts=OPTS_FIXED_WORK, profile=profile)
                assert ret_ == ITERS_PER_STEP and int(info_[5]) == ITERS_PER_STEP, (ret_, info_)
                if profile:
                    for kname, (ms, cnt) in B.profile().items():
                        a = prof.setdefault(kname, [0.0, 0]); a[0] += ms; a[1] += cnt
                last = info_
            return last
        e0.record(stream)
        info = run_steps(prof_in_timed)
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms_total = e0.elapsed_time(e1)
        c1 = B.counters()
        # second pass of the same steps: N = 1 without the event nodes and with attempts queued two ahead throughout; N > 1 with them
        barrier()
        g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
        B.reset(); B.solve(itmax=ITERS_PER_STEP, opts=OPTS_FIXED_WORK, profile=not prof_in_timed)
        g0.record(stream)
        run_steps(not prof_in_timed)
        g1.record(stream)
        barrier()
        ms_graph = g0.elapsed_time(g1)
        # time to convergence with the reference's default options, outside the timed region
        B.reset(); torch.cuda.synchronize()
        tc0 = time.perf_counter()
        retc, infoc = B.solve(itmax=150)
        conv = {"iterations": int(infoc[5]), "stop_reason": int(infoc[6]), "ms": 1e3 * (time.perf_counter() - tc0),
                "initial_cost": float(infoc[0]), "final_cost": float(infoc[1]), "opts": "default (1e-3, 1e-12, 1e-12, 1e-12, 0), itmax 150"}
    t_ms = torch.tensor([ms_total, ms_graph], dtype=torch.float64, device=device)
    tot_obs = torch.tensor([float(nobs_local)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX); dist.all_reduce(tot_obs, op=dist.ReduceOp.SUM)
    ms_total = float(t_ms[0].item()); ms_graph = float(t_ms[1].item()); total_obs = float(tot_obs.item())
    value = total_obs * ITERS_PER_STEP * args.steps / (ms_total * 1e-3)
    final_cost = float(info[1])

    # ---- end to end through the C ABI with HOST buffers: create + upload (H2D) + solve + download (D2H) + destroy ----
    e2e = None
    if not args.no_e2e:
        n_e2e = max(1, min(args.steps, 3))
        hps = [torch.from_numpy(p_l.copy()).pin_memory().numpy() for _ in range(n_e2e + 1)]  # p is in/out: one pinned copy per step (+ 1 warm-up)
        hx = torch.from_numpy(np.ascontiguousarray(x_l)).pin_memory().numpy()
        hv = torch.from_numpy(np.ascontiguousarray(vm_l).view(np.int8)).pin_memory().numpy()

        e2e_stage_ms = []          # N > 1, rank 0: [create + communicator, upload, solve, download, destroy] per call

        def e2e_step(hp):
            if world == 1:
                pe, ie, re_ = ba.solve_motstr(hv, hp, rot0, hx, nk, nd, itmax=ITERS_PER_STEP, opts=OPTS_FIXED_WORK, inplace=True)
            else:
                with torch.cuda.stream(stream):
                    tq = [time.perf_counter()]
                    Be = ba.BundleAdjuster(device=local, stream=stream.cuda_stream)
                    Be.share_nccl(B, rank, world)          # the process keeps ONE communicator; creating one is not part of a solve
                    Be.set_graph(False); tq.append(time.perf_counter())
                    Be.upload(hv, hp, rot0, hx, nk, nd); tq.append(time.perf_counter())
                    re_, ie = Be.solve(itmax=ITERS_PER_STEP, opts=OPTS_FIXED_WORK); tq.append(time.perf_counter())
                    Be.download(out=hp); tq.append(time.perf_counter())
                    Be.close(); tq.append(time.perf_counter())
                    e2e_stage_ms.append([round(1e3 * (b_ - a_), 2) for a_, b_ in zip(tq, tq[1:])])
            assert re_ == ITERS_PER_STEP
        e2e_step(hps[-1])                                  # one untimed call (first use of this path in the process)
        barrier()
        step_ms = []
        t0 = time.perf_counter()
        for hp in hps[:n_e2e]:
            ts = time.perf_counter()
            e2e_step(hp)
            step_ms.append(1e3 * (time.perf_counter() - ts))
        barrier()
        # per-call times vary by box (one call in a few takes 1.5-10 x the others: allocation of the multi-GB arena, host threads):
        # the value is taken from the MEDIAN call (max over ranks), the mean over all calls is reported beside it
        dt = torch.tensor([time.perf_counter() - t0, float(np.median(step_ms)) * 1e-3 * n_e2e], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt_mean, dt_med = float(dt[0].item()), float(dt[1].item())
        per_step_h2d = int(hv.nbytes + hx.nbytes + 8 * (16 * m + 3 * n_l) + 8 * 4 * m)
        e2e = {"value": total_obs * ITERS_PER_STEP * n_e2e / dt_med, "unit": "obs*iters/s",
               "value_mean_of_calls": total_obs * ITERS_PER_STEP * n_e2e / dt_mean, "value_is": "median call (max over ranks)",
               "h2d_bytes_per_step": per_step_h2d, "d2h_bytes_per_step": int(8 * (16 * m + 3 * n_l)), "steps": n_e2e,
               "step_ms_rank0": [round(v, 2) for v in step_ms], "stage_ms_rank0": e2e_stage_ms[1:], "bytes": "per rank", "api": "argus_ba_motstr_kdrts (host pointers: alloc + index build + H2D + %d LM iterations + D2H)" % ITERS_PER_STEP
               if world == 1 else "argus_ba_create + set_comm_nccl (the process's communicator) + upload + solve + download + destroy per step, pinned host buffers"}

    # ---- roofline of the dominant kernel (CUDA events on the launching stream, inside the timed region) ----
    kern = {k_: v for k_, v in prof.items() if v[1] > 0 and k_ not in ("allreduce", "residual", "linearize_stats")}
    dom = max(kern, key=lambda k_: kern[k_][0] / kern[k_][1])
    dom_ms = kern[dom][0] / kern[dom][1]
    sum_ms = sum(v[0] / v[1] for v in kern.values()) * kern[dom][1]
    v_l = vm_l.sum(axis=1).astype(np.float64)
    schur_flops_local = 768.0 * float((v_l * (v_l + 1.0)).sum())
    if dom == "schur":
        rl = {"kernel": dom, "bound": "tensor", "achieved": schur_flops_local / (dom_ms * 1e-3) / 1e12, "peak": dmma_peak,
              "unit": "TFLOP/s", "peak_source": "fp64 tensor pipe (DMMA m8n8k4), " + fp64_src}
    else:
        per = {"linearize": 40.0 * nobs_local + 72.0 * n_l, "backsub_eval": 40.0 * nobs_local + 72.0 * n_l}.get(dom, 40.0 * nobs_local)
        rl = {"kernel": dom, "bound": "hbm", "achieved": per / (dom_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src}
    rl["frac"] = rl["achieved"] / rl["peak"]
    rl["traffic"] = None
    try:        # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture of the same command
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if tr.get("workload") == args.workload and world == 1 and dom in tr:
            rl["traffic"] = tr[dom]; rl["traffic_unit"] = "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)"
    except Exception:
        pass
    rl["ms_per_launch"] = dom_ms
    rl["events_from"] = "the timed region (event-record nodes in every fourth attempt's graph)" if world == 1 else "a second pass of the same steps right after the timed region"
    rl["share_of_kernel_time"] = kern[dom][0] / sum_ms
    t_iter = ms_total * 1e-3 / (ITERS_PER_STEP * args.steps)
    flops_iter_job = counts["flops_iter"] * (1 if strong else world)
    bytes_iter_job = counts["bytes_iter"] * (1 if strong else world)
    rl["iteration"] = {"ms": t_iter * 1e3, "fp64_frac": flops_iter_job / t_iter / 1e12 / (dmma_peak * world),
                       "hbm_frac": bytes_iter_job / t_iter / 1e9 / (hbm_peak * world),
                       "algorithmic_gflop": flops_iter_job / 1e9, "algorithmic_gbyte": bytes_iter_job / 1e9,
                       ("ms_graph_replay" if world == 1 else "ms_profiled_pass"): ms_graph / (ITERS_PER_STEP * args.steps)}
    # one sample per phase and step: event-record nodes inside the captured attempt hold the times of the step's last attempt
    rl["kernels_ms_per_iteration"] = {k_: (v[0] / v[1] if v[1] else 0.0) for k_, v in sorted(prof.items()) if k_ not in ("residual", "linearize_stats")}
    rl["once_per_solve_ms"] = {k_: v[0] / v[1] for k_, v in sorted(prof.items()) if v[1] and k_ in ("residual", "linearize_stats")}
    rl["fp64_peaks_tflops"] = {"dmma_m8n8k4": dmma_peak, "dfma": dfma_peak, "source": fp64_src}

    # ---- CPU baseline: the reference's own C code on the host cores of this box (rank 0, N = 1) ----
    cpu = None
    if rank == 0 and world == 1:
        from oracle import sba_ref, ba_port
        vm, xs, ps = reference_sample(s, p0, m, args.cpu_baseline_points)
        kind = "reference" if sba_ref.available() else "port"
        fn = sba_ref.ref_motstr if kind == "reference" else ba_port.port_motstr
        t0 = time.perf_counter()
        _, iref, _ = fn(ps, xs, vm, rot0, nk, nd, itmax=ITERS_PER_STEP, opts=OPTS_FIXED_WORK)
        dt_ref = time.perf_counter() - t0
        assert int(iref[5]) == ITERS_PER_STEP, iref
        cpu = {"value": int(vm.sum()) * ITERS_PER_STEP / dt_ref, "unit": "obs*iters/s", "cores": 1, "kind": kind, "host_cores": os.cpu_count(),
               "sample": "first %d points (%d observations) of the same scene, %d cameras, %d LM iterations, %.1f s"
                         % (vm.shape[0], int(vm.sum()), m, ITERS_PER_STEP, dt_ref)}

    B.close()
    bd_res = None
    if not args.no_blockdiag and args.workload != "c5" and world == 1:
        bd_res = bench_blockdiag(torch, ba, stream, local, s, p0, rot0, m, nk, nd, max(2, min(args.steps, 5)),
                                 args.cpu_baseline_points if (rank == 0 and world == 1) else 0)
    pair_res = bench_pair(rank) if (not args.no_blockdiag and world == 1) else None
    del s
    extras = None
    if world == 1 and rank == 0 and not args.no_extras:
        extras = bench_extras(torch, ba, local, stream, args, dmma_peak)
    dlt_res = None
    if not args.no_dlt:
        dlt_res = bench_dlt(torch, device, stream, max(3, min(args.steps, 10)), 3, rank, world, hbm_peak, dfma_peak)
        if world > 1:
            v = torch.tensor([dlt_res["points_per_sec"]], dtype=torch.float64, device=device)
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
            dlt_res["points_per_sec"] = float(v.item())

    if rank == 0:
        if strong:
            wl = "%s: BA %d cameras, %d points x %d views = %d observations%s, 16 free parameters per camera" % (
                args.workload, m, n, views, int(total_obs), (" sharded over %d GPUs (balanced by observations)" % world) if world > 1 else "")
        else:
            wl = "%s: BA %d cameras, %d points x %d views = %d observations per GPU, 16 free parameters per camera" % (args.workload, m, n, views, nobs_local)
        line = {"metric": "ba_observations_per_second", "value": value, "unit": "obs*iters/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl,
                           "iters_per_step": ITERS_PER_STEP, "parallelism": "points sharded x%d, cameras replicated" % world,
                           "l2": "inputs larger than L2 (>= %.1f GB of per-observation blocks per pass and rank)" % (nobs_local * 48 * 8 / 1e9),
                           "final_cost": final_cost},
                "convergence": conv,
                "lm_iters_per_sec": ITERS_PER_STEP * args.steps / (ms_total * 1e-3),
                ("graph_replay" if world == 1 else "profiled_pass"):
                    {"value": total_obs * ITERS_PER_STEP * args.steps / (ms_graph * 1e-3), "unit": "obs*iters/s",
                     "note": ("same steps without the event-record nodes (the timed region carries them in every fourth attempt and waits for the "
                              "outcome of the attempts before it)") if world == 1 else
                             ("same steps with event-record nodes in every fourth attempt: the source of roofline.kernels_ms_per_iteration at "
                              "N > 1, where the timed region runs without them")},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(c1["launches"] - c0["launches"]),
                "allreduces": int(c1["allreduces"] - c0["allreduces"]), "roofline": rl, "cpu_baseline": cpu, "blockdiag": bd_res, "pair_init": pair_res,
                "configs": extras, "dlt": dlt_res}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
