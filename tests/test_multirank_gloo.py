"""CPU, world_size 2, gloo: the multi-rank plumbing of the BA path - block sharding of the points, the all-reduce hook
contract (sum / max on a raw buffer, in place) and the algebra it relies on: summing the per-shard partial normal
equations [U | ea | Schur sums | E] over ranks gives the unsharded reduced camera system (checked with the numpy oracle,
the GPU kernels are exercised by tests/test_gpu_ba.py::test_m0_two_shards_equal_single)."""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from argus_gui_b200 import scenes, ba
    from oracle import ba_port, ba_numpy
    m, n, mu = 5, 400, 0.25
    s = scenes.make_ba_scene(m, n, 3, seed=42)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    vm, p_loc, x_loc, (lo, hi) = ba.shard_problem(s["vmask"], p0, s["x"], m, world, rank, balance_observations=True)
    nl = hi - lo
    A, B = ba_port.port_jacobian(p_loc, vm, rot0, 2, 3)
    hx = ba_port.port_project(p_loc, vm, rot0)
    pi, ci = np.nonzero(vm)
    ne = ba_numpy.normal_equations(A, B, x_loc - hx, pi, ci, nl, m, mu)
    # what a rank contributes: U, ea, and the Schur parts  (S = blockdiag(U + mu I) - Schur,  E = ea - Eschur)
    Ublk = np.zeros((16 * m, 16 * m))
    for j in range(m):
        Ublk[16 * j:16 * j + 16, 16 * j:16 * j + 16] = ne["U"][j] + mu * np.eye(16)
    schur = Ublk - ne["S"]
    esch = ne["ea"].ravel() - ne["E"]
    buf = np.ascontiguousarray(np.concatenate([ne["U"].ravel(), ne["ea"].ravel(), schur.ravel(), esch, [float(((x_loc - hx) ** 2).sum())]]))
    hook = ba.host_allreduce_hook()
    assert hook(buf.ctypes.data, buf.size, 0, 0) == 0                       # sum, in place
    mx = np.array([np.abs(ne["eb"]).max(), float(rank)])
    assert hook(mx.ctypes.data, mx.size, 1, 0) == 0                          # max
    assert mx[1] == world - 1
    if rank == 0:
        np.savez(os.path.join(out_dir, "reduced.npz"), buf=buf, mx=mx)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_equals_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = np.load(str(tmp_path / "reduced.npz"))
    from argus_gui_b200 import scenes
    from oracle import ba_port, ba_numpy
    m, n, mu = 5, 400, 0.25
    s = scenes.make_ba_scene(m, n, 3, seed=42)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    A, B = ba_port.port_jacobian(p0, s["vmask"], rot0, 2, 3)
    hx = ba_port.port_project(p0, s["vmask"], rot0)
    pi, ci = np.nonzero(s["vmask"])
    ne = ba_numpy.normal_equations(A, B, s["x"] - hx, pi, ci, n, m, mu)
    buf = r["buf"]
    o = 0
    U = buf[o:o + m * 256].reshape(m, 16, 16); o += m * 256
    ea = buf[o:o + m * 16].reshape(m, 16); o += m * 16
    schur = buf[o:o + (16 * m) ** 2].reshape(16 * m, 16 * m); o += (16 * m) ** 2
    esch = buf[o:o + 16 * m]; o += 16 * m
    cost = buf[o]
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(U, ne["U"]) < 1e-13 and rel(ea, ne["ea"]) < 1e-12
    Ublk = np.zeros((16 * m, 16 * m))
    for j in range(m):
        Ublk[16 * j:16 * j + 16, 16 * j:16 * j + 16] = U[j] + mu * np.eye(16)
    assert rel(Ublk - schur, ne["S"]) < 1e-12
    assert rel(ea.ravel() - esch, ne["E"]) < 1e-11
    assert abs(cost - ((s["x"] - hx) ** 2).sum()) < 1e-9 * cost
    assert abs(r["mx"][0] - np.abs(ne["eb"]).max()) < 1e-12 * r["mx"][0]
