"""Drop-in for tiddit/DBSCAN.py (reference lines cited per function), computed on the GPU.

Same names, argument order and return types as the reference: float64 label arrays with -1 for
noise and the reference's own cluster ids.  Coordinates must be integers in [0, 2^31-1) (TIDDIT
passes genome positions); anything else raises instead of silently computing something different.
"""
import numpy as np

from . import _lib, device_ops

__all__ = ["main", "x_coordinate_clustering", "y_coordinate_clustering"]


def _column(data, c):
    arr = np.asarray(data)
    if arr.ndim == 1 and arr.size == 0:
        return np.zeros(0, dtype=np.int32)
    if arr.ndim != 2 or arr.shape[1] <= c:
        raise IndexError("too many indices for array")  # what data[i,:] / point[c] would raise
    col = arr[:, c]
    if col.dtype.kind == "f":
        if not np.all(col == np.floor(col)):
            raise TypeError("tiddit_b200.DBSCAN works on integer coordinates")
    elif col.dtype.kind not in "iu":
        raise TypeError("tiddit_b200.DBSCAN works on integer coordinates")
    if col.size and (col.min() < 0 or col.max() >= device_ops.INT32_MAX):
        raise OverflowError("coordinates must lie in [0, 2^31-1)")
    return np.ascontiguousarray(col, dtype=np.int32)


def x_coordinate_clustering(data, epsilon, m):
    """DBSCAN.py:33-64 -> (float64 labels, cluster_id).  Second caller: tiddit_contig_analysis.pyx:176."""
    torch = _lib.torch_cuda()
    x = _column(data, 0)
    device_ops.check_min_pts(m, len(x))
    if len(x) == 0:
        return np.zeros(0), -1
    labels, last = device_ops.xpass_device(torch.from_numpy(x).cuda(), epsilon, m)
    return labels.cpu().numpy().astype(np.float64), int(last.item())


def y_coordinate_clustering(data, epsilon, m, cluster_id, clusters):
    """DBSCAN.py:66-123 -> (clusters, cluster_id); `clusters` is rewritten in place like the reference.

    Ids are visited in ascending order, which is what iterating set(clusters) does for the dense ids
    0..cluster_id the x-pass produces."""
    torch = _lib.torch_cuda()
    y = _column(data, 1)
    device_ops.check_min_pts(m, len(y))
    if len(y) == 0:
        return clusters, cluster_id
    lab = np.asarray(clusters)
    if len(lab) != len(y):
        raise IndexError("boolean index did not match indexed array")
    lab32 = np.ascontiguousarray(lab, dtype=np.int32)
    lab_d = torch.from_numpy(lab32).cuda()
    cid_d = torch.tensor([int(cluster_id)], dtype=torch.int32, device="cuda")
    device_ops.ypass_device(torch.from_numpy(y).cuda(), epsilon, m, lab_d, cid_d)
    clusters[...] = lab_d.cpu().numpy()
    return clusters, int(cid_d.item())


def main(data, epsilon, m):
    """DBSCAN.py:125-129 -> float64 labels (the x-pass then the y-pass, ids as the reference numbers them)."""
    torch = _lib.torch_cuda()
    x = _column(data, 0)
    device_ops.check_min_pts(m, len(x))
    if len(x) == 0:
        return np.zeros(0)
    y = _column(data, 1)
    labels = device_ops.dbscan_main_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), epsilon, m)
    return labels.cpu().numpy().astype(np.float64)
