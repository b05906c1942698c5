"""TEST INFRASTRUCTURE ONLY (oracle shim) -- never imported by the product path.

Stand-in for torchsparse v1.4.0 ``sparse_quantize`` (reference README.md:27; call site
models/cluster.py:11).  Algorithm restated from the copy the reference vendors at
dataset_toolbox/prep_nuscene_waymo_sf/libs/spv_utils.py:7-20,57-60,81: floor(coords/voxel)
-> int32, ravel hash (subtract per-dim min, uint64 Horner with max+1 radices), np.unique with
first-occurrence indices ordered by ascending hash.  Parity unpinned by the reference (no tests).
"""
import numpy as np


def ravel_hash(x):
    assert x.ndim == 2
    x = x - np.min(x, axis=0)
    x = x.astype(np.uint64, copy=False)
    xmax = np.max(x, axis=0).astype(np.uint64) + 1
    h = np.zeros(x.shape[0], dtype=np.uint64)
    for k in range(x.shape[1] - 1):
        h += x[:, k]
        h *= xmax[k + 1]
    h += x[:, -1]
    return h


def sparse_quantize(coords, voxel_size=1, *, return_index=False, return_inverse=False):
    if isinstance(voxel_size, (float, int)):
        voxel_size = tuple(voxel_size for _ in range(3))
    voxel_size = np.array(voxel_size)
    coords = np.floor(coords / voxel_size).astype(np.int32)
    _, indices, inverse_indices = np.unique(ravel_hash(coords), return_index=True, return_inverse=True)
    coords = coords[indices]
    outputs = [coords]
    if return_index:
        outputs += [indices]
    if return_inverse:
        outputs += [inverse_indices]
    return outputs[0] if len(outputs) == 1 else outputs
