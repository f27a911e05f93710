"""Headless Clicker save path (argus_gui_b200/clicker_io.py): file formats on CPU against a loop-for-loop restatement of the
reference's helpers, the reconstruction behind them on the GPU against the DLT oracle."""
import numpy as np
import pandas
import pytest

from argus_gui_b200 import clicker_io, scenes
from oracle import clicker_port, dlt_port


def test_dense_to_sparse_and_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    arr = rng.normal(0, 100, (37, 16)); arr[rng.random(arr.shape) < 0.4] = np.nan; arr[5] = np.nan
    trip = clicker_io.dense_to_sparse(arr)
    assert np.array_equal(trip, clicker_port.dense_to_sparse(arr))
    assert np.array_equal(clicker_io.sparse_to_dense(trip), arr, equal_nan=True)
    p = str(tmp_path / "x-xypts.tsv")
    clicker_io.write_sparse(p, arr)
    head = open(p).readline().strip().split("\t")
    assert head == ["Row", "Column", "Entry"]
    back = clicker_io.load_tsv(p)
    assert np.array_equal(back, arr, equal_nan=True)
    # the reference's reader (float16, zeros for missing entries) sees the same numbers
    ref = clicker_port.load_tsv(p)
    assert ref.shape == arr.shape and np.array_equal(ref, np.nan_to_num(arr, nan=0.0).astype(np.float16))
    empty = np.full((4, 6), np.nan)
    assert np.array_equal(clicker_io.dense_to_sparse(empty), clicker_port.dense_to_sparse(empty))


@pytest.mark.gpu
@pytest.mark.parametrize("sparse", [True, False])
def test_save_tracks_files(tmp_path, sparse):
    d = scenes.make_dlt_scene(150, m=3, ntracks=2, seed=4, nan_frac=0.2)
    prefix = str(tmp_path / "proj")
    res = clicker_io.save_tracks(prefix, d["pts"], [list(p) for p in d["prof"]], d["dlt"], sparse=sparse, bootstrap=True, bs_iter=30, seed=2)
    if sparse:
        xyz = clicker_io.load_tsv(prefix + "-xyzpts.tsv"); err = clicker_io.load_tsv(prefix + "-xyzres.tsv")
        xy = clicker_io.load_tsv(prefix + "-xypts.tsv"); ci = clicker_io.load_tsv(prefix + "-xyz-cis.tsv")
        assert clicker_io.load_tsv(prefix + "-spline-weights.tsv").shape == res["weights"].shape
        assert clicker_io.load_tsv(prefix + "-spline-error-tolerances.tsv").shape == (1, 6)
    else:
        rd = lambda p: pandas.read_csv(p, float_precision="round_trip")
        fx = rd(prefix + "-xyzpts.csv"); fr = rd(prefix + "-xyzres.csv"); fp = rd(prefix + "-xypts.csv")
        assert list(fx.columns) == ["Track 1_x", "Track 1_y", "Track 1_z", "Track 2_x", "Track 2_y", "Track 2_z"]
        assert list(fr.columns) == ["Track 1", "Track 2"] and fp.columns[0] == "Track 1_cam_1_x" and fp.columns[-1] == "Track 2_cam_3_y"
        fc = rd(prefix + "-xyz-cis.csv")
        assert list(fc.columns[:6]) == ["Track 1_x_lower", "Track 1_y_lower", "Track 1_z_lower", "Track 1_x_upper", "Track 1_y_upper", "Track 1_z_upper"]
        xyz, err, xy, ci = fx.values, fr.values, fp.values, fc.values
    assert np.array_equal(xy, d["pts"], equal_nan=True)
    for t in range(2):
        xo = dlt_port.uv_to_xyz(d["pts"][:, 6 * t:6 * t + 6], d["prof"], d["dlt"])
        ok = ~np.isnan(xo)
        assert np.array_equal(np.isnan(xyz[:, 3 * t:3 * t + 3]), ~ok) and np.abs(xyz[:, 3 * t:3 * t + 3][ok] - xo[ok]).max() <= 1e-9 * np.abs(xo[ok]).max()
    eo = dlt_port.get_repo_errors(res["xyz"], d["pts"], d["prof"], d["dlt"]).T
    ok = ~np.isnan(eo)
    assert np.array_equal(np.isnan(err), ~ok) and np.abs(err[ok] - eo[ok]).max() <= 1e-9 * np.abs(eo[ok]).max()
    # lower < xyz < upper wherever a confidence interval exists
    lo, up = ci[:, 0:3], ci[:, 3:6]
    okc = ~np.isnan(lo).any(axis=1) & ~np.isnan(res["xyz"][:, :3]).any(axis=1)
    assert okc.any() and (lo[okc] <= res["xyz"][okc, :3]).all() and (res["xyz"][okc, :3] <= up[okc]).all()
