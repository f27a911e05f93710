"""Headless version of the Clicker save path ("on-disk formats"): what
/root/reference/argus_gui/resources/scripts/argus-click does when the user saves a project with DLT coefficients and camera
profiles loaded (save_sparse :3234-3335 and its dense twin) without the GUI around it:

    xypts  -> uv_to_xyz per track -> xyzpts, get_repo_errors -> xyzres, optional bootstrapXYZs -> xyz-cis, spline weights,
    tolerances; sparse "Row / Column / Entry" TSVs (1-based, first data row = [rows, columns, entries], :3146-3166) or dense
    CSVs; load_tsv (:1229-1247) reads the sparse files back.

The numerical work is one fused GPU call for all tracks (argus_dlt_reconstruct) instead of the reference's per-track,
per-frame Python loops; the file formats are the reference's.
"""
import numpy as np
import pandas

from . import dlt as _dlt
from . import tools

__all__ = ["dense_to_sparse", "sparse_to_dense", "load_tsv", "write_sparse", "reconstruct_tracks", "save_tracks"]


def dense_to_sparse(arr):
    """(rows, columns, entries) header row followed by 1-based (row, column, value) triplets of the non-NaN entries in
    row-major order (argus-click:3146-3166)."""
    arr = np.asarray(arr, dtype=np.float64)
    r, c = np.nonzero(~np.isnan(arr))
    out = np.zeros((len(r) + 1, 3))
    out[0] = (arr.shape[0], arr.shape[1], len(r))
    out[1:, 0] = r + 1
    out[1:, 1] = c + 1
    out[1:, 2] = arr[r, c]
    return out


def sparse_to_dense(trip, fill=np.nan):
    trip = np.asarray(trip, dtype=np.float64)
    out = np.full((int(trip[0, 0]), int(trip[0, 1])), fill)
    out[trip[1:, 0].astype(np.int64) - 1, trip[1:, 1].astype(np.int64) - 1] = trip[1:, 2]
    return out


def write_sparse(path, arr):
    df = pandas.DataFrame(dense_to_sparse(arr), columns=['Row', 'Column', 'Entry'])
    df[['Row', 'Column']] = df[['Row', 'Column']].astype(np.uint32)
    df.to_csv(path, index=False, na_rep='NaN', sep='\t')


def load_tsv(path, dtype=np.float64):
    """Dense array from a sparse TSV written by the Clicker (argus-click:1229-1247: the first data row carries the shape;
    missing entries are 0 there - it builds a sparse matrix - and NaN here, pass ``fill`` through sparse_to_dense for zeros).
    The reference reads into float16; ``dtype`` keeps that choice with the caller."""
    data = pandas.read_csv(path, index_col=False, sep='\t', float_precision='round_trip').values
    return sparse_to_dense(data).astype(dtype)


def reconstruct_tracks(pts, camera_profile, dlt_coefficients, bootstrap=False, bs_iter=250, seed=None):
    """pts: frames x (2 * cameras * tracks) pixel coordinates (NaN = not digitised) ->
    dict(xyz (frames x 3 tracks), res (frames x tracks), and with ``bootstrap`` cis (frames x 6 tracks: lower | upper per
    track, argus-click:3311-3320), weights, tols)."""
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    m = len(camera_profile)
    if pts.ndim != 2 or pts.shape[1] % (2 * m) != 0:
        raise tools.ArgusError("pts must be frames x (2 * cameras * tracks)")
    xyz, err = _dlt.reconstruct(pts, camera_profile, np.asarray(dlt_coefficients, dtype=np.float64))
    out = {"xyz": xyz, "res": err.T.copy()}
    if bootstrap:
        CIs, weights, tols = tools.bootstrapXYZs(pts, out["res"], camera_profile, dlt_coefficients, bsIter=bs_iter, seed=seed)
        upper, lower = xyz + CIs, xyz - CIs
        k = xyz.shape[1] // 3
        ci = np.full((xyz.shape[0], 6 * k), np.nan)
        for t in range(k):
            ci[:, 6 * t:6 * t + 6] = np.concatenate((lower[:, 3 * t:3 * t + 3], upper[:, 3 * t:3 * t + 3]), axis=1)
        out.update(cis=ci, weights=weights, tols=np.asarray(tols, dtype=np.float64))
    return out


def save_tracks(prefix, pts, camera_profile=None, dlt_coefficients=None, track_names=None, sparse=True, bootstrap=False, bs_iter=250, seed=None):
    """Write the Clicker's output files for ``pts`` under ``prefix``: ``-xypts``, and when profiles and DLT coefficients are
    given ``-xyzpts``, ``-xyzres`` (+ ``-xyz-cis``, ``-spline-weights``, ``-spline-error-tolerances`` with ``bootstrap``);
    ``.tsv`` (sparse) or ``.csv`` (dense, with the reference's column names).  Returns the dict of reconstruct_tracks (or {})."""
    pts = np.asarray(pts, dtype=np.float64)
    res = {}
    m = len(camera_profile) if camera_profile is not None else None
    if m is not None and dlt_coefficients is not None:
        # the bootstrap's sub-frame search shifts its input in place (tools.bootstrapXYZs, as the reference's does); the Clicker
        # builds the -xypts table BEFORE it bootstraps (argus-click:3280 vs :3311), so the file holds the digitised points
        work = pts.copy() if bootstrap else pts
        res = reconstruct_tracks(work, camera_profile, dlt_coefficients, bootstrap=bootstrap, bs_iter=bs_iter, seed=seed)
    k = pts.shape[1] // (2 * m) if m else None
    names = list(track_names) if track_names is not None else (['Track %d' % (t + 1) for t in range(k)] if k else None)
    if sparse:
        write_sparse(prefix + '-xypts.tsv', pts)
        if res:
            write_sparse(prefix + '-xyzpts.tsv', res["xyz"])
            write_sparse(prefix + '-xyzres.tsv', res["res"])
            if bootstrap:
                write_sparse(prefix + '-xyz-cis.tsv', res["cis"])
                write_sparse(prefix + '-spline-weights.tsv', res["weights"])
                write_sparse(prefix + '-spline-error-tolerances.tsv', res["tols"].reshape(1, -1))
    else:
        if names and m:
            cols = ['%s_cam_%d_%s' % (n, c + 1, a) for n in names for c in range(m) for a in ('x', 'y')]
        else:
            cols = None
        pandas.DataFrame(pts, columns=cols).to_csv(prefix + '-xypts.csv', index=False, na_rep='NaN')
        if res:
            xc = ['%s_%s' % (n, a) for n in names for a in ('x', 'y', 'z')]
            pandas.DataFrame(res["xyz"], columns=xc).to_csv(prefix + '-xyzpts.csv', index=False, na_rep='NaN')
            pandas.DataFrame(res["res"], columns=names).to_csv(prefix + '-xyzres.csv', index=False, na_rep='NaN')
            if bootstrap:
                cc = ['%s_%s_%s' % (n, a, b) for n in names for b in ('lower', 'upper') for a in ('x', 'y', 'z')]
                pandas.DataFrame(res["cis"], columns=cc).to_csv(prefix + '-xyz-cis.csv', index=False, na_rep='NaN')
    return res
