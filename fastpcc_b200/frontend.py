"""Device front end of the codec path: float scan -> unique Morton-sorted voxels, and the kd-tree partition of large
clouds.  Mirrors the host code the reference runs right before `Model.compress` / `compress_partitions`:
lib/datasets/KITTIOdometry/dataset.py:90-102,117-118 and lib/data_utils.py:95-113,163-234 (`kd_tree_partition`,
as called by `pc_data_collate_fn`).  Everything runs on the device through the C ABI (fpcc_voxelize_f32,
fpcc_kd_split); the only host reads are the voxel count and the three split integers per kd node, which decide the
recursion exactly as the reference's Python recursion does."""
from typing import List, Tuple

import torch

from . import _lib
from .ops import _p, _s, workspace


def _call(name, *args):
    _lib.call(name, *args)


def voxelize(points: torch.Tensor, resolution: int = 65536, span: float = 400.0, batch: int = 0,
             morton_inverse: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """points: float32 CUDA [N, >=3] (KITTI .bin rows are x,y,z,reflectance).  Returns (coords int32 [M,4] =
    (batch,x,y,z) unique voxels in Morton order, inv_transform float32 [4] = (origin xyz, 400/(resolution-1))), the
    `xyz` and `inv_transform` the reference dataset hands to the model (dataset.py:90-102,117-120)."""
    if not points.is_cuda or points.dtype != torch.float32 or points.dim() != 2 or points.shape[1] < 3:
        raise RuntimeError('voxelize: expected a float32 CUDA tensor [N, >=3]')
    if points.stride(1) != 1:
        points = points.contiguous()
    n, dev = points.shape[0], points.device
    if n == 0:
        raise RuntimeError('voxelize: empty scan')
    bits = max(1, int(resolution - 1).bit_length())
    out = torch.empty((n, 4), dtype=torch.int32, device=dev)
    inv = torch.empty(4, dtype=torch.float32, device=dev)
    cnt = torch.empty(1, dtype=torch.int32, device=dev)
    ws = workspace(_lib.load().fpcc_voxelize_workspace(n), dev)
    scale = (resolution - 1) / span
    _call('fpcc_voxelize_f32', _p(points), n, points.stride(0), scale, bits, 0 if morton_inverse else 2, batch,
          _p(out), _p(inv), _p(cnt), _p(ws), ws.numel(), _s())
    inv[3] = span / (resolution - 1)
    m = int(cnt.item())
    if m < 0:
        raise RuntimeError(f'voxelize: a quantised coordinate does not fit {bits} bits (resolution {resolution})')
    return out[:m], inv


def kd_split(coords: torch.Tensor, coord_bits: int):
    """One kd node: (left rows, right rows, axis, split value); rows keep their input order."""
    n, dev = coords.shape[0], coords.device
    out = torch.empty_like(coords)
    info = torch.empty(3, dtype=torch.int32, device=dev)
    ws = workspace(_lib.load().fpcc_kd_split_workspace(n, coord_bits), dev)
    _call('fpcc_kd_split', _p(coords), n, coord_bits, _p(out), _p(info), _p(ws), ws.numel(), _s())
    axis, value, n_left = info.tolist()
    return out[:n_left], out[n_left:], axis, value


def kd_tree_partition(coords: torch.Tensor, max_num: int, coord_bits: int = None) -> List[torch.Tensor]:
    """`kd_tree_partition(xyz, max_num)` of lib/data_utils.py:163-234 for coordinates (int32 CUDA [N,4], batch
    column first as after the collate's pad, data_utils.py:141-143): same partitions, same order, rows in input order."""
    if not coords.is_cuda or coords.dtype != torch.int32 or coords.dim() != 2 or coords.shape[1] != 4:
        raise RuntimeError('kd_tree_partition: expected an int32 CUDA tensor [N,4]')
    coords = coords.contiguous()
    if coord_bits is None:
        coord_bits = max(1, int(coords[:, 1:].max().item()).bit_length())
    n = coords.shape[0]
    if n <= max_num:
        return [coords]
    left, right, _, _ = kd_split(coords, coord_bits)
    if n // 2 <= max_num:
        return [left.clone(), right.clone()]
    return kd_tree_partition(left.clone(), max_num, coord_bits) + kd_tree_partition(right.clone(), max_num, coord_bits)


def collate_partitions(coords: torch.Tensor, max_num: int) -> List[torch.Tensor]:
    """The `xyz` list `pc_data_collate_fn` builds when a cloud exceeds `kd_tree_partition_max_points_num`
    (data_utils.py:127-143): element 0 is the whole cloud, the partitions follow.  Feed to `compress_partitions`."""
    if coords.shape[0] <= max_num:
        return [coords]
    return [coords] + kd_tree_partition(coords, max_num)
