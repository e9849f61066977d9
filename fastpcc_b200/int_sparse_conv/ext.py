"""Drop-in for the reference's pybind11 module `int_sparse_conv_ext`
(lib/int_sparse_conv/src/binding.cu:114-145): same function / class names, same argument meaning, same
error behaviour (RuntimeError where the reference has TORCH_CHECK), all forwarding to the C ABI."""
import torch

from .. import ops

_DIVISOR = 128  # hashmap_cuda.cuh:102


def _check_cuda(*ts):
    dev = ts[0].device
    if not dev.type == 'cuda':
        raise RuntimeError('Expected device.is_cuda()')
    for t in ts:
        if t.device != dev:
            raise RuntimeError('Expected all tensors on the same CUDA device')
        if not t.is_contiguous():
            raise RuntimeError('Expected contiguous tensors')


def cutlass_gemm_int8(A, B, C, D):
    """gemm.cu:140-247: D = A @ B.T + C, C is (N,), (M,N) or empty."""
    _check_cuda(A, B, C, D)
    ops.gemm_i8(A, B, C, D)


def cutlass_gather_gemm_scatter_int8(A, B, C, D, gather_idx, scatter_idx):
    """gather_gemm_scatter.cu:157-268: D[scatter[i]] = C[scatter[i]] + A[gather[i]] @ B.T."""
    _check_cuda(A, B, C, D, gather_idx, scatter_idx)
    L = gather_idx.numel()
    if scatter_idx.numel() != L or L > D.size(0):
        raise RuntimeError('Expected scatter_idx.size(0) == L && L <= D_rows')
    if C.data_ptr() != D.data_ptr():
        idx = scatter_idx.long()
        if C.numel() == 0:
            D[idx] = 0
        elif C.dim() == 1:
            D[idx] = C[None]
        else:
            D[idx] = C[idx]
    ops.gather_gemm_scatter_i8(A, B, D, gather_idx, scatter_idx)


def softmax_int32(input):
    if not input.is_cuda or not input.is_contiguous():
        raise RuntimeError('Expected a contiguous CUDA tensor')
    return ops.softmax_i32(input)


def _rq(out_type, input, requant_mul, zero_point, shift, bias=None, slope=None):
    ts = [input, requant_mul, zero_point] + [t for t in (bias, slope) if t is not None]
    _check_cuda(*ts)
    if input.dim() != 2 or requant_mul.dim() != 1:
        raise RuntimeError('Expected input.dim() == 2 && requant_mul.dim() == 1')
    ch = input.size(1)
    if not (input.size(0) > 0 and ch > 0 and shift >= 0):
        raise RuntimeError('Expected N > 0 && Ch > 0 && right_shift >= 0')
    if requant_mul.size(0) != ch or zero_point.numel() != 1 or (bias is not None and bias.size(0) != ch) \
            or (slope is not None and slope.numel() != 1):
        raise RuntimeError('Expected Ch == bias.size(0) && 1 == slope.numel() && Ch == requant_mul.size(0) && 1 == zero_point.numel()')
    if ch == 1:  # a [1] multiplier would otherwise be read as "scalar"; identical result
        pass
    return ops.requant(input, ops.make_epilogue(requant_mul, zero_point, shift, out_type, bias=bias, slope=slope))


def requant_to_int8(input, requant_mul, zero_point, shift): return _rq(ops.OUT_I8, input, requant_mul, zero_point, shift)
def requant_to_int16(input, requant_mul, zero_point, shift): return _rq(ops.OUT_I16, input, requant_mul, zero_point, shift)
def requant_to_int32(input, requant_mul, zero_point, shift): return _rq(ops.OUT_I32, input, requant_mul, zero_point, shift)
def bias_requant_to_int8(input, bias, requant_mul, zero_point, shift): return _rq(ops.OUT_I8, input, requant_mul, zero_point, shift, bias=bias)
def bias_requant_to_int16(input, bias, requant_mul, zero_point, shift): return _rq(ops.OUT_I16, input, requant_mul, zero_point, shift, bias=bias)
def bias_requant_to_int32(input, bias, requant_mul, zero_point, shift): return _rq(ops.OUT_I32, input, requant_mul, zero_point, shift, bias=bias)
def prelu_requant_to_int8(input, slope, requant_mul, zero_point, shift): return _rq(ops.OUT_I8, input, requant_mul, zero_point, shift, slope=slope)
def prelu_requant_to_int16(input, slope, requant_mul, zero_point, shift): return _rq(ops.OUT_I16, input, requant_mul, zero_point, shift, slope=slope)
def prelu_requant_to_int32(input, slope, requant_mul, zero_point, shift): return _rq(ops.OUT_I32, input, requant_mul, zero_point, shift, slope=slope)
def bias_prelu_requant_to_int8(input, bias, slope, requant_mul, zero_point, shift): return _rq(ops.OUT_I8, input, requant_mul, zero_point, shift, bias=bias, slope=slope)
def bias_prelu_requant_to_int16(input, bias, slope, requant_mul, zero_point, shift): return _rq(ops.OUT_I16, input, requant_mul, zero_point, shift, bias=bias, slope=slope)
def bias_prelu_requant_to_int32(input, bias, slope, requant_mul, zero_point, shift): return _rq(ops.OUT_I32, input, requant_mul, zero_point, shift, bias=bias, slope=slope)


def prelu(input, slope):
    _check_cuda(input, slope)
    if input.dim() != 2 or slope.numel() != 1 or input.numel() == 0:
        raise RuntimeError('Expected input.dim() == 2 && N > 0 && Ch > 0 && 1 == slope.numel()')
    return ops.prelu_i32(input, slope)


class GPUHashTable:
    """hashmap_cuda.cuh:87-145.  Non-owning view over caller tensors (keys int64, vals int32), or an owning
    table of `capacity` slots.  Coordinates arrive in the reference's permuted (x,y,z,batch) order."""

    def __init__(self, keys_or_capacity, vals=None):
        if vals is None:
            cap = int(keys_or_capacity)
            self.keys = torch.zeros(cap, dtype=torch.int64, device='cuda')
            self.vals = torch.zeros(cap, dtype=torch.int32, device='cuda')
        else:
            self.keys, self.vals = keys_or_capacity, vals
            if self.keys.dtype != torch.int64 or self.vals.dtype != torch.int32:
                raise RuntimeError('GPUHashTable expects int64 keys and int32 vals')

    def insert_coords(self, coords):
        ops.hash_build(coords, layout=1, keys=self.keys, vals=self.vals)

    def lookup_coords(self, coords, kernel_sizes, strides, kernel_volume):
        ks = [int(v) for v in kernel_sizes.tolist()]
        st = [int(v) for v in strides.tolist()]
        if ks[0] * ks[1] * ks[2] != kernel_volume:
            raise RuntimeError('kernel_volume does not match kernel_sizes')
        n = coords.size(0)
        rows = (n + _DIVISOR - 1) // _DIVISOR * _DIVISOR
        return ops.kmap_lookup(self.keys, self.vals, coords, ks, st, layout=1, k_major=False, pad_rows=rows)

    def insert_vals(self, keys):
        raise NotImplementedError('generic key->value insert is not on the codec path (never called by cuda_ops.py)')

    def lookup_vals(self, keys):
        raise NotImplementedError('generic key->value lookup is not on the codec path (never called by cuda_ops.py)')
