"""Occupancy / residual heads of the lossy object codecs on the GPU coders
(models/convolutional/lossy_coord_lossy_color/geo_lossl_em.py:59-114 of the reference):

  init_prob ................ P(1) = clip(round(sigmoid(logit) * 65536), 1, 65535)                  (:95-99)
  binary_encode / decode ... BinaryRansCoder over those probabilities, batch 1                      (:101-114)
  rans_encode_with_cdf / rans_decode_with_cdf ... histogram-CDF coding of rounded residuals with the
      stream layout  u24 count | [u8 -offset] | u8 len(cdf)-2 | u16 cdf[1:-1] | u24 bytes | payload  (:59-93)

Probabilities are derived from FLOAT logits, so encoder and decoder must see identical bits: the conv kernels
accumulate in a fixed order and the sigmoid/round below are evaluated by the same device code on both sides.
"""
import io
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib, ops
from .ops import _p, _s
from .rans_coder import BinaryRansCoder, IndexedRansCoder


def int_to_bytes(x: int, length: int) -> bytes:
    return int(x).to_bytes(length, 'little', signed=False)


def bytes_to_int(s: bytes) -> int:
    return int.from_bytes(s, 'little', signed=False)


def init_prob(dist: torch.Tensor) -> torch.Tensor:
    """-> int32 [n] on the device (uint32 bit patterns of P(1) * 65536 in [1, 65535])"""
    p = torch.round(dist.sigmoid().double() * 65536.0)  # sigmoid in the logits' own dtype, as the reference evaluates it;  # np.round and torch.round both round half to even
    return p.clamp_(1, 65535).to(torch.int32).reshape(-1)


def binary_encode(dist: torch.Tensor, x: torch.Tensor) -> bytes:
    """one stream over all symbols (batch_size == 1 in the reference)"""
    assert dist.shape[0] == x.shape[0]
    n = x.numel()
    prob = init_prob(dist)
    sym = x.reshape(-1).to(torch.uint8).contiguous()
    ranges = torch.empty(n, dtype=torch.int32, device=prob.device)
    _lib.call('fpcc_rans_binary_ranges', _p(sym), _p(prob), n, _p(ranges), _s())
    off = torch.tensor([0, n], dtype=torch.int64, device=prob.device)
    cap = (2 * n + 64 + 3) & ~3
    out, out_len = ops.rans_encode(ranges, off, cap)
    size = int(out_len.item())
    assert size > 0
    return out[0, cap - size:].cpu().numpy().tobytes()


def binary_decode(dist: torch.Tensor, data: bytes) -> torch.Tensor:
    prob = init_prob(dist)
    n = prob.numel()
    dev = prob.device
    blob = torch.frombuffer(bytearray(data) + bytearray(ops.RansDecodeStreams.PAD), dtype=torch.uint8).to(dev)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    off = torch.zeros(1, dtype=torch.int64, device=dev)
    ln = torch.tensor([len(data)], dtype=torch.int32, device=dev)
    _lib.call('fpcc_rans_binary_decode', _p(blob), _p(off), _p(ln), _p(prob), n, 1, _p(out), _p(err), _s())
    if int(err.item()):
        raise RuntimeError('binary_decode: read past the end of the stream')
    return out.bool().reshape(dist.shape[0])


def rans_encode_with_cdf(target: np.ndarray, bs: io.BytesIO, offset: Optional[int] = None, shape_bytes: int = 3):
    coder = IndexedRansCoder(False, 1)
    bs.write(int_to_bytes(target.shape[0], shape_bytes))
    if offset is None:
        offset = int(target.min())
        bs.write(int_to_bytes(-offset, 1))
    pmf = np.bincount((target - offset).reshape(-1)).astype(np.float64)
    coder.init_with_pmfs(pmf[None], np.array([offset], dtype=np.int32))
    cdf = coder.get_cdfs()[0]
    bs.write(int_to_bytes(len(cdf) - 2, 1))
    for cd in cdf[1:-1]:
        bs.write(int_to_bytes(cd, 2))
    payload = coder.encode(np.ascontiguousarray(target.reshape(1, -1), dtype=np.int32))[0]
    bs.write(int_to_bytes(len(payload), 3))
    bs.write(payload)


def rans_decode_with_cdf(bs: io.BytesIO, channels: int, offset: Optional[int] = None, shape_bytes: int = 3) -> Tuple[np.ndarray, List[int]]:
    coder = IndexedRansCoder(False, 1)
    shape_sum = bytes_to_int(bs.read(shape_bytes))
    if offset is None:
        offset = -bytes_to_int(bs.read(1))
    cdf = [0, *(bytes_to_int(bs.read(2)) for _ in range(bytes_to_int(bs.read(1)))), 1 << 16]
    coder.init_with_quantized_cdfs([cdf], np.array([offset], dtype=np.int32))
    payload = bs.read(bytes_to_int(bs.read(3)))
    target = np.empty((1, shape_sum * channels), np.int32)
    coder.decode([payload], target)
    return target.reshape(shape_sum, channels), cdf


@torch.no_grad()
def get_keep(pred_f: torch.Tensor, coords: torch.Tensor, tensor_stride, max_stride, target_points_num=None) -> torch.Tensor:
    """Top-k pruning mask of the lossy decoder (lossy_coord_v2/layers.py:151-180), device side.

    pred_f [n] or [n,1]: occupancy logits of the candidate voxels `coords` [n,4] = (batch, x, y, z) of stride
    `tensor_stride`; candidates are grouped by their ancestor voxel of stride `max_stride` (MinkowskiMaxPooling +
    MinkowskiPoolingTranspose with kernel = stride = max_stride / tensor_stride): the best candidate of every group
    is always kept; among the others, sample b keeps those above its (n_b - target_b)-th smallest logit
    (`torch.kthvalue`), or above 0 when no target counts are given."""
    f = pred_f.reshape(-1)
    n = f.shape[0]
    ts = torch.as_tensor(tensor_stride, device=coords.device, dtype=torch.int64)
    ms = torch.as_tensor(max_stride, device=coords.device, dtype=torch.int64)
    assert bool((ms % ts == 0).all())
    c = coords.long()
    anc = torch.div(c[:, 1:], ms, rounding_mode='floor')  # ancestor voxel index (coordinates are multiples of tensor_stride)
    lo = anc.amin(0)
    span = (anc.amax(0) - lo + 1)
    key = ((c[:, 0] * span[0] + (anc[:, 0] - lo[0])) * span[1] + (anc[:, 1] - lo[1])) * span[2] + (anc[:, 2] - lo[2])
    _, inv = torch.unique(key, return_inverse=True)
    gmax = torch.full((int(inv.max().item()) + 1,), float('-inf'), dtype=f.dtype, device=f.device)
    gmax.scatter_reduce_(0, inv, f, reduce='amax')
    not_max = (f - gmax[inv]) != 0
    if target_points_num is not None:
        thr = torch.empty(len(target_points_num), dtype=f.dtype, device=f.device)
        for b, tgt in enumerate(target_points_num):
            rows = c[:, 0] == b
            nb = int(rows.sum().item())
            assert nb > tgt
            thr[b] = torch.kthvalue(f[rows & not_max], nb - tgt, dim=0).values
        threshold = thr[c[:, 0]]
    else:
        threshold = 0
    keep = f > threshold
    keep |= ~not_max
    return keep
