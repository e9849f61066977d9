"""Multi-threaded CPU backend of oracle/int_ops.py for the TIMED CPU arm of bench.py (`cpu_baseline`,
`--impl reference`): the same operators with the same integer results, evaluated the way a competent CPU
implementation of the reference path would -- GEMMs on the BLAS of torch-CPU (fp32 SGEMM over all host cores, exact
because every partial sum of an int8 x int8 contraction of K <= 1040 stays below 2^24), the element-wise epilogues
as fused in-place int64 tensor ops (OpenMP inside ATen), the per-offset gather / scatter of the sparse conv as
index_select / index_add_ -- instead of single-threaded numpy passes with float64 / int64 temporaries.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned to oracle/int_ops.py function by function on random
inputs and to the reference's golden bitstreams through the whole codec (tests/test_oracle_fast_backend.py).

Usage:   from oracle import lossl_coord_int, int_ops_fast;  lossl_coord_int.K = int_ops_fast
Every name of int_ops that is not overridden here is re-exported unchanged.
"""
import ctypes as C
import os.path as osp
import subprocess

import numpy as np
import torch

from .int_ops import *  # noqa: F401,F403  (kernel_offsets, lookup_coords, compact_kernel_map, exp_lut, ...)
from .int_ops import SharedFxpShift, EXP_LUT_SIZE, _INT_RANGE, exp_lut, lookup_coords, compact_kernel_map, _wrap32  # noqa: F401

_TDT = {np.int8: torch.int8, np.int16: torch.int16, np.int32: torch.int32}
_OUT_CODE = {np.int8: 0, np.int16: 1, np.int32: 2}
_HERE = osp.dirname(osp.abspath(__file__))
_SO = osp.join(_HERE, '_build', 'libcpu_kernels.so')
_ck = None


def ck():
    """libcpu_kernels.so (oracle/cpu_kernels.c, OpenMP): one fused pass per element-wise operator"""
    global _ck
    if _ck is None:
        if not osp.isfile(_SO) or osp.getmtime(_SO) < osp.getmtime(osp.join(_HERE, 'cpu_kernels.c')):
            subprocess.run(['make', '-C', _HERE], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(_SO)
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
        L.fo_requant.argtypes = [vp, i64, C.c_int, vp, C.c_int, i32, vp, i64, C.c_int, C.c_int, vp]
        L.fo_prelu.argtypes = [vp, i64, i32, vp]
        L.fo_quantize_pmf.argtypes = [vp, i64, C.c_int, i64, vp, C.c_int, vp]
        L.fo_f32_to_i32_add.argtypes = [vp, i64, C.c_int, vp, C.c_int, vp]
        L.fo_f32_to_i32_add.restype = C.c_double
        L.fo_scatter_add_f32.argtypes = [vp, i64, C.c_int, vp, vp]
        L.fo_gather_i8_to_f32.argtypes = [vp, C.c_int, vp, i64, vp]
        _ck = L
    return _ck


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def set_threads(n):
    torch.set_num_threads(int(n))


def _t(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        return torch.from_numpy(a.view(np.int32)).to(torch.int64) & 0xFFFFFFFF
    return torch.from_numpy(a)


def _rha_(x: torch.Tensor, shift: int) -> torch.Tensor:
    """round-half-away-from-zero arithmetic right shift, in place on int64 (requant.cu:16-20):
    (x + half - [x < 0]) >> shift  ==  sign(x) * ((|x| + half) >> shift)."""
    if shift <= 0:
        return x
    neg = (x < 0).to(torch.int64)
    x.add_((1 << (shift - 1))).sub_(neg)
    x.bitwise_right_shift_(shift)
    return x


def _prelu64_(v: torch.Tensor, slope: int) -> torch.Tensor:
    scaled = _rha_(v * slope, 25)
    return torch.where(v < 0, scaled, v)


def requant(inp, requant_mul, zero_point, shift, out_dtype, bias=None, slope=None):
    assert inp.dtype == np.int32 and inp.ndim == 2 and shift >= 0
    inp = np.ascontiguousarray(inp)
    n, ch = inp.shape
    mul = np.ascontiguousarray(np.asarray(requant_mul).astype(np.uint32, copy=False).reshape(-1))
    assert mul.shape[0] == ch
    b = None if bias is None else np.ascontiguousarray(np.asarray(bias).astype(np.int32, copy=False).reshape(-1))
    out = np.empty((n, ch), dtype=out_dtype)
    ck().fo_requant(_ptr(inp), n, ch, None if b is None else _ptr(b), 0 if slope is None else 1,
                    0 if slope is None else int(np.asarray(slope).reshape(-1)[0]), _ptr(mul),
                    int(np.asarray(zero_point).reshape(-1)[0]), int(shift), _OUT_CODE[out_dtype], _ptr(out))
    return out


def prelu(inp, slope):
    assert inp.dtype == np.int32
    inp = np.ascontiguousarray(inp)
    out = np.empty_like(inp)
    ck().fo_prelu(_ptr(inp), inp.size, int(np.asarray(slope).reshape(-1)[0]), _ptr(out))
    return out


def _mm(a_i8: torch.Tensor, b_i8: torch.Tensor) -> torch.Tensor:
    """a [M,K] int8 @ b [N,K].T exactly -> int32 (fp32 SGEMM while K*128*128 < 2^24, else fp64)"""
    K = a_i8.shape[1]
    ft = torch.float32 if K * 128 * 128 < (1 << 24) else torch.float64
    return (a_i8.to(ft) @ b_i8.to(ft).T).to(torch.int32)


_wcache = {}


def _wt(w: np.ndarray) -> torch.Tensor:
    """float image of a weight block, converted once (weights are static)"""
    key = (w.__array_interface__['data'][0], w.shape, w.strides)
    t = _wcache.get(key)
    if t is None:
        K = w.shape[-1]
        ft = torch.float32 if K * 128 * 128 < (1 << 24) else torch.float64
        t = torch.from_numpy(np.ascontiguousarray(w)).to(ft)
        if len(_wcache) > 4096:
            _wcache.clear()
        _wcache[key] = (t, w)  # keep the array alive: the key is its address
        return t
    return t[0]


def _a_f32(A, gather=None):
    """int8 rows (optionally gathered) -> float32 matrix for the BLAS GEMM, one parallel pass"""
    A = np.ascontiguousarray(A)
    L = A.shape[0] if gather is None else gather.shape[0]
    out = torch.empty((L, A.shape[1]), dtype=torch.float32)
    ck().fo_gather_i8_to_f32(_ptr(A), A.shape[1], None if gather is None else _ptr(gather), L, out.data_ptr())
    return out


def gemm_int8(A, B, C=None):
    """cutlass_gemm_int8 (gemm.cu): D = A @ B.T + C, int32"""
    assert A.dtype == np.int8 and B.dtype == np.int8
    wt = _wt(B)
    if wt.dtype != torch.float32:  # K beyond fp32 exactness: float64 path
        d = torch.from_numpy(np.ascontiguousarray(A)).to(wt.dtype) @ wt.T
        if C is not None and C.size:
            c = torch.from_numpy(np.ascontiguousarray(C)).to(d.dtype)
            d.add_(c[None] if c.dim() == 1 else c)
        assert (float(d.abs().max()) if d.numel() else 0.0) < 2 ** 31, 'int32 accumulator overflow'
        return d.to(torch.int32).numpy()
    acc = _a_f32(A) @ wt.T
    out = np.empty(acc.shape, dtype=np.int32)
    mode, add = 0, None
    if C is not None and C.size:
        add = np.ascontiguousarray(C.astype(np.int32, copy=False))
        mode = 1 if add.ndim == 1 else 2
    worst = ck().fo_f32_to_i32_add(acc.data_ptr(), acc.shape[0], acc.shape[1], None if add is None else _ptr(add), mode, _ptr(out))
    assert worst < 2 ** 31, 'int32 accumulator overflow'
    return out


def gather_gemm_scatter_int8(A, B, D, gather_idx, scatter_idx):
    """D[scatter[i]] += A[gather[i]] @ B.T in place (cuda_ops.py:163-166); scatter indices are unique within an offset"""
    wt = _wt(B)
    g = np.ascontiguousarray(gather_idx).astype(np.int32, copy=False)
    s = np.ascontiguousarray(scatter_idx).astype(np.int32, copy=False)
    assert wt.dtype == torch.float32 and D.dtype == np.int32 and D.flags.c_contiguous
    upd = _a_f32(A, g) @ wt.T
    ck().fo_scatter_add_f32(upd.data_ptr(), g.shape[0], upd.shape[1], _ptr(s), _ptr(D))


def sparse_conv_in8w8out32(in_feats, weight, in_coords, out_coords, kernel_size, stride,
                           in_out_maps=None, zero_point_comp=None, if_in_coords_equals_out_coords=False):
    """cuda_ops.py:95-169 (same control flow as int_ops.sparse_conv_in8w8out32)"""
    kv = int(np.prod(kernel_size))
    idx_omit = kv >> 1 if (if_in_coords_equals_out_coords and all(k % 2 == 1 for k in kernel_size)) else -1
    if in_out_maps is None:
        table = lookup_coords(in_coords, out_coords, kernel_size, stride)
        in_out_maps = compact_kernel_map(table, idx_omit)
    in_feats = np.ascontiguousarray(in_feats)
    out = np.zeros((out_coords.shape[0], weight.shape[1]), dtype=np.int32)
    if idx_omit != -1:
        c = out if zero_point_comp is None else out + zero_point_comp[idx_omit][None]
        out = gemm_int8(in_feats, weight[idx_omit], c)
    for k, (im, om) in enumerate(in_out_maps):
        if im is not None:
            if zero_point_comp is not None:
                out[om] += zero_point_comp[k][None]
            gather_gemm_scatter_int8(in_feats, weight[k], out, im, om)
    return out, in_out_maps


_lut_t = None


def softmax_int32(x: np.ndarray) -> np.ndarray:
    """softmax.cu:41-106"""
    global _lut_t
    assert x.dtype == np.int32 and x.ndim == 2
    if _lut_t is None:
        _lut_t = torch.from_numpy(exp_lut().astype(np.int64))
    x64 = torch.from_numpy(np.ascontiguousarray(x)).to(torch.int64)
    row_max = x64.amax(1, keepdim=True) + 64
    idx = ((row_max - x64) >> 7).clamp_(max=EXP_LUT_SIZE - 1)
    e = _lut_t[idx]
    row_sum = e.sum(1, keepdim=True)
    inv = torch.where(row_sum > 0, torch.div((1 << 32) + (row_sum >> 1), row_sum.clamp(min=1), rounding_mode='floor'),
                      torch.full_like(row_sum, (1 << 32) // x.shape[1]))
    prod = e * inv  # < 2^16 * 2^33: fits int64
    return prod.clamp_(max=0xFFFFFFFF).numpy().astype(np.uint32)


def batch_quantize_pmf(logits: np.ndarray) -> np.ndarray:
    """lossl_coord_int/model.py:344-353 over softmax.cu:41-106, one fused pass per row"""
    assert logits.dtype == np.int32 and logits.ndim == 2
    if logits.strides[1] != 4 or logits.strides[0] % 4:
        logits = np.ascontiguousarray(logits)
    n, S = logits.shape
    lut = exp_lut()
    out = np.empty((n, S), dtype=np.uint16)
    ck().fo_quantize_pmf(_ptr(logits), n, S, logits.strides[0] // 4 if n > 1 else S, _ptr(lut), lut.shape[0], _ptr(out))
    return out
