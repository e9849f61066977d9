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
