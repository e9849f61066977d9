"""numpy restatement of the reference's host front end: voxelisation of a float scan and the kd-tree partition.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity status: `kd_tree_partition` is PINNED to the
reference's own lib/data_utils.py (imported unmodified by tests/golden/make_frontend_golden.py; fixture
tests/golden/frontend_golden.json).  `voxelize` restates five NumPy lines of a dataset method that cannot be called
without the KITTI files (lib/datasets/KITTIOdometry/dataset.py:90-102,117-118): pinned to the source.
"""
import numpy as np

from .lossl_coord_int import morton_xmajor


def voxelize(points_f32: np.ndarray, resolution: int = 65536, span: float = 400.0, morton: bool = True):
    """dataset.py:90-102: scale = (resolution-1)/400; org = xyz.min(0); xyz -= org; xyz *= scale (float32, in place);
    round half to even; int32; np.unique(axis=0); then the Morton argsort of dataset.py:117-118 with inverse=True
    (x most significant), the order compress() establishes anyway (lossl_coord_int/model.py:398).
    Returns (xyz int32 [M,3], org float32 [3], inv_scale)."""
    assert points_f32.dtype == np.float32
    xyz = points_f32[:, :3].copy()
    scale = (resolution - 1) / span
    org = xyz.min(0)
    xyz -= org
    xyz *= scale
    q = np.unique(xyz.round().astype(np.int32), axis=0)
    if morton:
        q = q[np.argsort(morton_xmajor(q), kind='stable')]
    return q, org, span / (resolution - 1)


def kd_tree_partition(coord: np.ndarray, max_num: int):
    """_kd_tree_partition (lib/data_utils.py:187-234) for coordinates only: split on the axis of largest variance at
    the (n//2)-th smallest value (ties go left), recurse while half the node still exceeds max_num."""
    n = len(coord)
    if n <= max_num:
        return [coord]
    dim = int(np.argmax(np.var(coord, 0)))
    k = n // 2
    split_value = np.partition(coord[:, dim], k - 1)[k - 1]  # torch.kthvalue(k): k-th smallest, 1-based
    mask = coord[:, dim] <= split_value
    if k <= max_num:
        return [coord[mask], coord[~mask]]
    return kd_tree_partition(coord[mask], max_num) + kd_tree_partition(coord[~mask], max_num)
