"""numpy restatement of the reference's integer sparse-conv operators (lib/int_sparse_conv).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Parity status: the reference ships NO golden vectors for these ops and its CUDA extension cannot
be built here (needs a GPU, torchsparse and the author's CUTLASS fork), so this restatement is
"parity unpinned" against reference OUTPUTS.  It is pinned against the reference SOURCE instead:
every function cites the lines it follows, the arithmetic is exact integer arithmetic whose result
does not depend on evaluation order, and tests/test_oracle_int_ops.py checks (a) the exp LUT against
the SHA-256 of the 6145-entry literal in src/softmax.cu:13-22 and (b) hand-computed cases for the
rounding / saturation rules.

Conventions follow the reference: coords are int32 [N,4] = (batch, x, y, z); weights are
[kernel_volume, C_out, C_in] int8; "Q8.23" is the shared fixed-point activation format.
"""
import math

import numpy as np

SharedFxpShift = 23  # cuda_ops.py:15
EXP_LUT_SIZE = 12 * 512 + 1  # softmax.cu:5

_INT_RANGE = {np.int8: (-128, 127), np.int16: (-32768, 32767), np.int32: (-2 ** 31, 2 ** 31 - 1)}


# ------------------------------------------------------------------------------------------------
# element-wise epilogues: src/element_wise/{requant,bias_requant,prelu_requant,bias_prelu_requant,prelu}.cu
# ------------------------------------------------------------------------------------------------

def rha_shift(x: np.ndarray, shift: int) -> np.ndarray:
    """round-half-away-from-zero arithmetic right shift on int64 (requant.cu:16-20); identity if shift==0."""
    x = x.astype(np.int64, copy=False)
    if shift <= 0:
        return x
    half = np.int64(1) << np.int64(shift - 1)
    pos = (x + half) >> np.int64(shift)
    neg = -((-x + half) >> np.int64(shift))
    return np.where(x >= 0, pos, neg)


def _prelu64(v: np.ndarray, slope: int) -> np.ndarray:
    """Q6.25 PReLU on int64 values (bias_prelu_requant.cu:17-22)."""
    scaled = rha_shift(v * np.int64(slope), 25)
    return np.where(v < 0, scaled, v)


def requant(inp, requant_mul, zero_point, shift, out_dtype, bias=None, slope=None):
    """{,bias_,prelu_,bias_prelu_}requant_to_int{8,16,32} (bias_prelu_requant.cu:6-37, requant.cu:7-26).

    inp int32 [N,Ch]; requant_mul uint32 [Ch]; zero_point int64 scalar; shift >= 0 (host int).
    """
    assert inp.dtype == np.int32 and inp.ndim == 2
    assert shift >= 0
    mul = np.asarray(requant_mul).astype(np.int64).reshape(1, -1)
    assert mul.shape[1] == inp.shape[1]
    v = inp.astype(np.int64)
    if bias is not None:
        v = v + np.asarray(bias).astype(np.int64).reshape(1, -1)
    if slope is not None:
        v = _prelu64(v, int(np.asarray(slope).reshape(-1)[0]))
    prod = v * mul + np.int64(int(np.asarray(zero_point).reshape(-1)[0]))
    out = rha_shift(prod, shift)
    lo, hi = _INT_RANGE[out_dtype]
    return np.clip(out, lo, hi).astype(out_dtype)


def prelu(inp, slope):
    """prelu (prelu.cu:6-21): Q8.23 in/out, Q6.25 slope, saturate to int32."""
    assert inp.dtype == np.int32
    v = _prelu64(inp.astype(np.int64), int(np.asarray(slope).reshape(-1)[0]))
    return np.clip(v, -2 ** 31, 2 ** 31 - 1).astype(np.int32)


# ------------------------------------------------------------------------------------------------
# integer softmax + CDF head: src/softmax.cu, lossl_coord_int/model.py:344-353
# ------------------------------------------------------------------------------------------------

_LUT = None


def exp_lut() -> np.ndarray:
    """lut[k] = llround(exp(-k/512) * 65536), k = 0..6144 (softmax.cu:108-117 `build_lut`; the literal at
    softmax.cu:13-22 holds the same values -- checked by SHA-256 in tests/test_oracle_int_ops.py)."""
    global _LUT
    if _LUT is None:
        _LUT = np.array([int(math.floor(math.exp(-k / 512.0) * 65536.0 + 0.5)) for k in range(EXP_LUT_SIZE)],
                        dtype=np.int32)
    return _LUT


def softmax_int32(x: np.ndarray) -> np.ndarray:
    """softmax_int32 (softmax.cu:41-106): int32 Q15.16 logits [N,C] -> uint32 Q0.32 probabilities."""
    assert x.dtype == np.int32 and x.ndim == 2
    lut = exp_lut().astype(np.int64)
    x64 = x.astype(np.int64)
    row_max = x64.max(1, keepdims=True) + 64  # "+ (1 << 6) for rounding", softmax.cu:71
    idx = np.minimum((row_max - x64) >> 7, EXP_LUT_SIZE - 1)
    e = lut[idx]
    row_sum = e.sum(1, keepdims=True)  # int32 in the kernel; 255*65536 cannot overflow
    inv = np.where(row_sum > 0,
                   ((np.int64(1) << 32) + (row_sum >> 1)) // np.maximum(row_sum, 1),
                   (np.int64(1) << 32) // x.shape[1])
    prod = e.astype(np.uint64) * inv.astype(np.uint64)
    return np.minimum(prod, np.uint64(0xFFFFFFFF)).astype(np.uint32)


def batch_quantize_pmf(logits: np.ndarray) -> np.ndarray:
    """Model.batch_quantize_pmf_torch (lossl_coord_int/model.py:344-353): Q8.23 logits [n,S] -> uint16
    inclusive CDF [n,S] whose last entry is forced to 65535."""
    assert logits.dtype == np.int32
    S = logits.shape[1]
    p = softmax_int32(logits >> (SharedFxpShift - 16)).astype(np.int64)
    pmf = ((p * (65536 - S)) >> 32) + 1
    cdf = np.cumsum(pmf, axis=1)
    cdf[:, -1] = 65535
    return cdf.astype(np.uint16)


# ------------------------------------------------------------------------------------------------
# kernel map: src/hashmap/hashmap_cuda.cuh:221-275 + cuda_ops.py:115-151
# ------------------------------------------------------------------------------------------------

def kernel_offsets(kernel_size):
    """Offset (dx,dy,dz) of every kernel index, in the enumeration order of lookup_coords_kernel
    (hashmap_cuda.cuh:239-258): odd kernel volume -> x fastest; even -> z fastest.  The per-axis
    offset is `k % ks - (ks-1)//2` in BOTH branches, i.e. 0..1 for ks=2 but -1..2 for ks=4."""
    ks = tuple(int(k) for k in kernel_size)
    vol = ks[0] * ks[1] * ks[2]
    out = np.zeros((vol, 3), dtype=np.int64)
    order = (0, 1, 2) if vol % 2 == 1 else (2, 1, 0)
    for kidx in range(vol):
        r = kidx
        for ax in order:
            out[kidx, ax] = r % ks[ax] - (ks[ax] - 1) // 2
            r //= ks[ax]
    return out


_OFF = 1 << 16
_BITS = 18


def _pack(bxyz: np.ndarray) -> np.ndarray:
    c = bxyz.astype(np.int64)
    xyz = c[:, 1:] + _OFF
    ok = ((xyz >= 0) & (xyz < (1 << _BITS))).all(1) & (c[:, 0] >= 0) & (c[:, 0] < 512)
    key = (c[:, 0] << (3 * _BITS)) | (xyz[:, 0] << (2 * _BITS)) | (xyz[:, 1] << _BITS) | xyz[:, 2]
    return np.where(ok, key, -1)


def lookup_coords(in_coords, out_coords, kernel_size, stride):
    """Dense neighbour table [kernel_volume, N_out]: value = input row index + 1, 0 = no neighbour
    (hashmap_cuda.cuh:221-275, transposed as at cuda_ops.py:129).  The reference keys its table by a
    64-bit FNV hash of the coordinate; this restatement matches exact coordinates (identical unless two
    coordinates collide in 64 bits)."""
    offs = kernel_offsets(kernel_size)
    K = offs.shape[0]
    n_out = out_coords.shape[0]
    in_keys = _pack(in_coords)
    assert (in_keys >= 0).all(), 'input coordinates out of the oracle key range'
    order = np.argsort(in_keys, kind='stable')
    sk = in_keys[order]
    assert (np.diff(sk) != 0).all(), 'duplicate input coordinates'
    oc = out_coords.astype(np.int64)
    st = np.asarray(stride, dtype=np.int64)
    table = np.zeros((K, n_out), dtype=np.int32)
    for k in range(K):
        q = oc.copy()
        q[:, 1:] = oc[:, 1:] * st[None] + offs[k][None]
        qk = _pack(q)
        pos = np.searchsorted(sk, qk)
        pos = np.minimum(pos, sk.shape[0] - 1)
        hit = (sk[pos] == qk) & (qk >= 0)
        table[k, hit] = order[pos[hit]].astype(np.int32) + 1
    return table


def compact_kernel_map(out_in_map, idx_omit_map=-1):
    """cuda_ops.py:132-151: per-offset (in_map, out_map) int32 pair lists, offset-major, output index
    ascending; (None, None) for empty offsets and for the omitted centre offset."""
    maps = []
    for k in range(out_in_map.shape[0]):
        if k == idx_omit_map:
            maps.append((None, None))
            continue
        out_map = np.nonzero(out_in_map[k])[0].astype(np.int32)
        if out_map.size == 0:
            maps.append((None, None))
        else:
            maps.append((out_in_map[k][out_map] - 1, out_map))
    return maps


# ------------------------------------------------------------------------------------------------
# GEMMs: src/gemm.cu:11-127, src/gather_gemm_scatter.cu:11-144
# ------------------------------------------------------------------------------------------------

def _exact_matmul(a_i8: np.ndarray, b_i8: np.ndarray) -> np.ndarray:
    """a [M,K] int8 @ b[N,K].T -> int64 exactly.  BLAS float GEMM is exact while every partial sum is
    an integer below the mantissa: 128*128*K < 2^24 for float32, < 2^53 for float64."""
    K = a_i8.shape[1]
    ft = np.float32 if K * 128 * 128 < (1 << 24) else np.float64
    return (a_i8.astype(ft) @ b_i8.astype(ft).T).astype(np.int64)


def _wrap32(x64: np.ndarray) -> np.ndarray:
    return (x64 & 0xFFFFFFFF).astype(np.uint32).view(np.int32)


def gemm_int8(A, B, C=None):
    """cutlass_gemm_int8 (gemm.cu): D = A @ B.T + C; C is None/empty, (N,) bias, or (M,N).  int32 result.
    (The reference MMA saturates its int32 accumulator; asserting no overflow keeps both definitions equal.)"""
    assert A.dtype == np.int8 and B.dtype == np.int8
    d = _exact_matmul(A, B)
    if C is not None and C.size:
        d = d + (C.astype(np.int64)[None] if C.ndim == 1 else C.astype(np.int64))
    assert np.abs(d).max(initial=0) < 2 ** 31, 'int32 accumulator overflow'
    return d.astype(np.int32)


def gather_gemm_scatter_int8(A, B, D, gather_idx, scatter_idx):
    """cutlass_gather_gemm_scatter_int8 with C == D (cuda_ops.py:163-166):
    D[scatter[i]] += A[gather[i]] @ B.T, in place."""
    upd = _exact_matmul(A[gather_idx], B)
    acc = D[scatter_idx].astype(np.int64) + upd
    assert np.abs(acc).max(initial=0) < 2 ** 31
    D[scatter_idx] = acc.astype(np.int32)  # scatter indices are unique within one offset


def sparse_conv_in8w8out32(in_feats, weight, in_coords, out_coords, kernel_size, stride,
                           in_out_maps=None, zero_point_comp=None, if_in_coords_equals_out_coords=False):
    """sparse_conv_in8w8out32 (cuda_ops.py:95-169).  Returns (out int32 [N2,C2], in_out_maps)."""
    kv = int(np.prod(kernel_size))
    idx_omit = kv >> 1 if (if_in_coords_equals_out_coords and all(k % 2 == 1 for k in kernel_size)) else -1
    if in_out_maps is None:
        table = lookup_coords(in_coords, out_coords, kernel_size, stride)
        in_out_maps = compact_kernel_map(table, idx_omit)
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
