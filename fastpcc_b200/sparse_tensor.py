"""SparseTensor / TensorCache shim with exactly the fields the reference reads and writes on
torchsparse 2.1.0 objects (lib/int_sparse_conv/cuda_ops.py:54-57,86,324-361;
models/convolutional/lossl_coord_int/model.py:63-69,173-174,190,280-291)."""
from typing import Any, Dict, Optional, Tuple, Union

import torch


class TensorCache:
    def __init__(self):
        self.cmaps: Dict[Tuple[int, ...], Tuple[torch.Tensor, Any]] = {}
        self.kmaps: Dict[Tuple[Any, ...], Any] = {}
        self.hashmaps: Dict[Tuple[int, ...], Tuple[torch.Tensor, torch.Tensor]] = {}


def _triple(x):
    return tuple(int(v) for v in x) if isinstance(x, (tuple, list)) else (int(x),) * 3


class SparseTensor:
    def __init__(self, feats: torch.Tensor, coords: torch.Tensor, stride: Union[int, Tuple[int, ...]] = 1,
                 spatial_range: Optional[Tuple[int, ...]] = None):
        self.feats = feats
        self.coords = coords
        self.stride = _triple(stride)
        self.spatial_range = spatial_range
        self._caches = TensorCache()

    @property
    def F(self):
        return self.feats

    @F.setter
    def F(self, v):
        self.feats = v

    @property
    def C(self):
        return self.coords

    @C.setter
    def C(self, v):
        self.coords = v

    @property
    def s(self):
        return self.stride

    @s.setter
    def s(self, v):
        self.stride = _triple(v)

    def to(self, device, non_blocking=True):
        self.feats = self.feats.to(device, non_blocking=non_blocking)
        self.coords = self.coords.to(device, non_blocking=non_blocking)
        return self

    def cuda(self):
        return self.to('cuda')
