"""TEST INFRASTRUCTURE - loop-for-loop restatement of the Clicker's sparse file helpers
(/root/reference/argus_gui/resources/scripts/argus-click: dense_to_sparse :3146-3166, load_tsv :1229-1247); the script itself
needs pyglet / Qt windows and cannot be imported.  Only tests/ may import this module."""
import numpy as np
import pandas


def dense_to_sparse(arr):
    row, col, val = [], [], []
    for k in range(len(arr)):
        for j in range(arr.shape[1]):
            if not np.isnan(arr[k, j]):
                row.append(k + 1); col.append(j + 1); val.append(arr[k, j])
    out = np.zeros((len(row) + 1, 3))
    out[1:, 0] = np.array(row); out[1:, 1] = np.array(col); out[1:, 2] = np.array(val)
    out[0, 0] = len(arr); out[0, 1] = arr.shape[1]; out[0, 2] = len(row)
    return out


def load_tsv(tsv):
    """-> dense float16 array with zeros where nothing was stored (the reference then wraps it in a csc_matrix)."""
    csv_data = pandas.read_csv(tsv, index_col=False, sep='\t', skiprows=1)
    data = np.zeros((int(list(csv_data.keys())[0]), int(list(csv_data.keys())[1])), dtype=np.float16)
    csv_data = csv_data.values
    for k in range(len(csv_data)):
        data[int(csv_data[k, 0]) - 1, int(csv_data[k, 1]) - 1] = csv_data[k, 2]
    return data
