"""Integer-only layer API, drop-in for lib/int_sparse_conv/cuda_ops.py of the reference: same class and
function names, constructor signatures, buffers (state-dict keys) and numerics -- but every layer is ONE
fused sm_100a kernel (gather + int8 GEMM + bias/PReLU/requant epilogue [+ residual]) instead of a Python
loop over kernel offsets followed by a separate element-wise pass.

Kernel maps are cached in `SparseTensor._caches.kmaps[(stride, kernel_size, conv_stride)]['in_out_maps']`
as a `KernelMap` (dense k-major neighbour table living on the device); the reference's list-of-pairs form
is produced on demand (`KernelMap.as_pair_lists()`) and accepted as input.
"""
import math
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..sparse_tensor import SparseTensor
from . import ext as int_sparse_conv_ext  # noqa: F401  (same module-level name as the reference)

SharedFxpShift = 23  # Q8.23   (cuda_ops.py:15)
WeightRange = (1 << 7) - 1
ActRange = (1 << 7) - 1


GROUP_ROWS_MIN = 4096  # below this a conv is launch-bound and the sort does not pay


class KernelMap:
    """Neighbour table [kernel_volume, n_out] int32: input row + 1, 0 = none (k-major, device)."""

    def __init__(self, table: torch.Tensor, idx_omit_map: int = -1):
        self.table = table
        self.idx_omit_map = idx_omit_map
        self._pairs = None
        self._grouped = None

    def grouped(self):
        """(table regrouped by neighbour pattern, row permutation) for the tensor-core conv, built once per map"""
        if self._grouped is None:
            self._grouped = ops.group_rows(self.table)
        return self._grouped

    @classmethod
    def from_pair_lists(cls, in_out_maps, n_out, kernel_volume, n_in_equals_out_centre=-1, device=None):
        """Accepts the reference's `in_out_maps` (list of (in_map, out_map) or (None, None))."""
        table = torch.zeros((kernel_volume, n_out), dtype=torch.int32, device=device)
        for k, (im, om) in enumerate(in_out_maps):
            if im is not None:
                table[k, om.long()] = im + 1
        if n_in_equals_out_centre >= 0:  # the omitted centre offset of an odd stride-1 kernel is the identity
            table[n_in_equals_out_centre] = torch.arange(1, n_out + 1, dtype=torch.int32, device=device)
        return cls(table, n_in_equals_out_centre)

    def as_pair_lists(self, idx_omit_map=None) -> List[Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]]:
        """cuda_ops.py:132-151 output format (one host sync to slice the per-offset ranges)."""
        omit = self.idx_omit_map if idx_omit_map is None else idx_omit_map
        if self._pairs is None or self._pairs[0] != omit:
            in_map, out_map, offsets = ops.kmap_compact(self.table, omit)
            off = offsets.tolist()
            maps = []
            for k in range(self.table.shape[0]):
                maps.append((None, None) if off[k + 1] == off[k] else (in_map[off[k]: off[k + 1]], out_map[off[k]: off[k + 1]]))
            self._pairs = (omit, maps)
        return self._pairs[1]

    # list-like access so that code written against the reference's list keeps working
    def __iter__(self):
        return iter(self.as_pair_lists())

    def __len__(self):
        return self.table.shape[0]

    def __getitem__(self, i):
        return self.as_pair_lists()[i]


def _cached(mod, attr, key, make):
    """Derived parameter tensors cached on a module are picked up by other threads on other CUDA streams (concurrent
    coding groups): ONE thread builds them under ops.cache_lock, the kernels that fill them finish before the entry
    becomes visible, and an entry is only replaced when the parameters it was derived from changed (ops.cache_lock)."""
    cache = getattr(mod, attr, None)
    if cache is not None and cache[0] == key:
        return cache
    with ops.cache_lock:
        cache = getattr(mod, attr, None)
        if cache is None or cache[0] != key:
            cache = (key,) + tuple(make())
            ops.publish_ready()
            setattr(mod, attr, cache)
    return cache


def _as_kernel_map(in_out_maps, n_out, kernel_volume, idx_omit_map, device):
    if isinstance(in_out_maps, KernelMap):
        return in_out_maps
    return KernelMap.from_pair_lists(in_out_maps, n_out, kernel_volume, idx_omit_map, device)


def build_kernel_map(in_coords, out_coords, kernel_size, stride, hashmap_kv=None):
    if hashmap_kv is None:
        hashmap_kv = ops.hash_build(in_coords, layout=0)
    table = ops.kmap_lookup(hashmap_kv[0], hashmap_kv[1], out_coords, kernel_size, stride, layout=0, k_major=True)
    return KernelMap(table), hashmap_kv


def sparse_conv_in8w8out32(
    in_feats: torch.Tensor,             # N1 x C1, scaled int
    weight: torch.Tensor,               # kernel_volume x C2 x C1
    in_coords: torch.Tensor,            # N1 x 4 (batch_idx, x, y, z)
    out_coords: torch.Tensor,           # N2 x 4
    kernel_size: Tuple[int, int, int],
    stride: Tuple[int, int, int],
    in_out_maps=None,
    hashmap_kv: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
    zero_point_comp: Optional[torch.Tensor] = None,
    if_in_coords_equals_out_coords: bool = False,
    _epilogue=None,
):
    """cuda_ops.py:95-169.  Returns (out N2 x C2 int32, hashmap_kv, in_out_maps)."""
    kernel_volume = math.prod(kernel_size)
    idx_omit_map = kernel_volume >> 1 if (if_in_coords_equals_out_coords and all(k % 2 == 1 for k in kernel_size)) else -1
    if in_out_maps is None:
        kmap, hashmap_kv = build_kernel_map(in_coords, out_coords, kernel_size, stride, hashmap_kv)
        kmap.idx_omit_map = idx_omit_map
    else:
        kmap = _as_kernel_map(in_out_maps, out_coords.shape[0], kernel_volume, idx_omit_map, in_feats.device)
    ep = _epilogue if _epilogue is not None else ops.identity_epilogue(in_feats.device)
    kv, c_out, c_in = weight.shape
    if (zero_point_comp is None and kv > 1 and kmap.table.shape[1] >= GROUP_ROWS_MIN
            and ops.gemm_engine(c_in, c_out, kv) == 'tc'):
        table_p, perm = kmap.grouped()
        out = ops.spconv(in_feats, weight, table_p, ep, row_perm=perm)
    else:
        out = ops.spconv(in_feats, weight, kmap.table, ep, zp_comp=zero_point_comp)
    return out, hashmap_kv, kmap


class LoadSaveUint32RequantMul(nn.Module):
    """torch.save lacks uint32: requant_mul travels as int64 (cuda_ops.py:172-186)."""

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        key = prefix + 'requant_mul'
        destination[key] = destination[key].to(torch.int64)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        key = prefix + 'requant_mul'
        if key in state_dict:
            state_dict[key] = state_dict[key].to(torch.uint32)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


def _shift_of(mod) -> int:
    """requant_shift as a host int, cached (the reference calls .item() on every forward)."""
    v = getattr(mod, '_shift_cache', None)
    if v is None or v[0] is not mod.requant_shift or v[2] != mod.requant_shift._version:
        v = (mod.requant_shift, int(mod.requant_shift.item()), mod.requant_shift._version)
        mod._shift_cache = v
    return v[1]


def _fill_requant(mod, requant_mul: torch.Tensor, zero_point_out: Optional[torch.Tensor]):
    shift = torch.log2((1 << (32 - mod.requant_mul_guard_bits)) / requant_mul).min().floor()
    assert shift >= 0, shift
    mod.requant_shift[:] = shift
    mod.requant_mul[:] = (requant_mul * 2 ** mod.requant_shift.to(torch.float)).round().to(torch.int64).to(torch.uint32)
    if zero_point_out is not None:
        mod.zero_point_out[:] = zero_point_out
        mod.int_zero_point_out[:] = zero_point_out.to(torch.int64) << mod.requant_shift
    else:
        mod.zero_point_out[:] = 0
        mod.int_zero_point_out[:] = 0


class _AffineIn8(LoadSaveUint32RequantMul):
    """Buffers shared by the conv and linear layers (cuda_ops.py:194-206, 516-528)."""

    def _register(self, weight_shape, out_ch, with_prelu):
        self.register_buffer('weight', torch.zeros(weight_shape, dtype=torch.int8), persistent=True)
        self.register_buffer('bias', torch.zeros((out_ch,), dtype=torch.int32), persistent=True)
        if with_prelu:
            self.register_buffer('slope', torch.zeros((1,), dtype=torch.int32), persistent=True)
        self.register_buffer('requant_mul', torch.zeros((out_ch,), dtype=torch.uint32), persistent=True)
        self.register_buffer('requant_shift', torch.zeros((1,), dtype=torch.int32), persistent=True)
        self.register_buffer('int_zero_point_out', torch.zeros((1,), dtype=torch.int64), persistent=True)
        self.register_buffer('scale_in', torch.zeros((1,), dtype=torch.float32) - 1, persistent=True)
        self.register_buffer('zero_point_in', torch.zeros((1,), dtype=torch.float32), persistent=True)
        self.register_buffer('scale_weight', torch.zeros((out_ch,), dtype=torch.float32) - 1, persistent=True)
        self.register_buffer('scale_out', torch.zeros((1,), dtype=torch.float32) - 1, persistent=True)
        self.register_buffer('zero_point_out', torch.zeros((1,), dtype=torch.float32), persistent=True)

    def epilogue(self, with_bias: bool, residual=None, post_slope=None, row_bias=None, post_requant=None, aux_out=None):
        shift = _shift_of(self)
        if self.out_scaled_int:
            out_type = ops.OUT_I8
        else:
            out_type, shift = ops.OUT_I32, shift - SharedFxpShift
        return ops.make_epilogue(self.requant_mul, self.int_zero_point_out, shift, out_type,
                                 bias=self.bias if with_bias else None,
                                 slope=self.slope if self.with_prelu else None,
                                 residual=residual, post_slope=post_slope, row_bias=row_bias, post_requant=post_requant, aux_out=aux_out)

    def _import_scales(self, scale_in, scale_out):
        assert scale_in.dtype == torch.float32 and scale_in.numel() == 1
        if scale_in <= self.eps:
            print(f'Warning: {scale_in}')
            scale_in = scale_in.clip(min=self.eps)
        if self.out_scaled_int:
            assert scale_out.dtype == scale_in.dtype and scale_out.numel() == 1
            if scale_out <= self.eps:
                print(f'Warning: {scale_out}')
                scale_out = scale_out.clip(min=self.eps)
            self.scale_out[:] = scale_out
        else:
            assert scale_out is None
        self.scale_in[:] = scale_in
        return scale_in, scale_out

    def _import_slope(self, prelu):
        if self.with_prelu:
            assert prelu.weight.numel() == 1
            if prelu.weight.abs().item() > 63:
                print(f'Warning: abnormal prelu slope {prelu.weight.item()}.')
            self.slope[:] = (prelu.weight * (1 << 25)).round().to(torch.int32)  # Q6.25
        else:
            assert prelu is None


class SparseConvIn8Out8(_AffineIn8):
    def __init__(self, in_ch: int, out_ch: int, kernel_size: Tuple[int, int, int], stride: Tuple[int, int, int],
                 with_prelu: bool, out_scaled_int: bool, eps=None, requant_mul_guard_bits=None):
        super().__init__()
        kernel_volume = kernel_size[0] * kernel_size[1] * kernel_size[2]
        self._register((kernel_volume, out_ch, in_ch), out_ch, with_prelu)
        self.in_ch, self.out_ch = in_ch, out_ch
        self.kernel_volume, self.kernel_size, self.stride = kernel_volume, tuple(kernel_size), tuple(stride)
        self.with_prelu = with_prelu
        self.use_zero_point_in = False
        self.out_scaled_int = out_scaled_int
        self.eps = eps if eps is not None else torch.finfo(torch.float32).eps
        self.requant_mul_guard_bits = requant_mul_guard_bits if requant_mul_guard_bits is not None else 10

    @torch.no_grad()
    def import_parameters(self, scale_in, zero_point_in, scale_out, zero_point_out, conv, prelu):
        """PTQ import from a float conv with `.kernel [K,Cin,Cout]`, `.bias`, `.kernel_size`, `.stride`
        (cuda_ops.py:223-301)."""
        assert conv.kernel.dtype == torch.float32 and conv.bias is not None
        assert zero_point_in.dtype in (torch.int32, torch.int64) and zero_point_in.numel() == 1
        assert self.weight.size(0) == conv.kernel.size(0)
        assert tuple(self.kernel_size) == tuple(conv.kernel_size) and tuple(self.stride) == tuple(conv.stride)
        scale_in, scale_out = self._import_scales(scale_in, scale_out)
        scale_weight = conv.kernel.abs().amax(dim=(0, 1)) / WeightRange
        if (scale_weight <= self.eps).any():
            print(f'Warning: {scale_weight}')
            scale_weight = scale_weight.clip(min=self.eps)
        self.scale_weight[:] = scale_weight
        self.weight[...] = (conv.kernel.permute((0, 2, 1)) / scale_weight[None, :, None]).round() \
            .clip(-WeightRange, WeightRange).to(torch.int8)
        if zero_point_in != 0:
            self.use_zero_point_in = True
            self.zero_point_in[:] = zero_point_in
            self.register_buffer('int_zero_point_in_comp',
                                 torch.zeros((self.kernel_volume, self.out_ch), dtype=torch.int32, device=self.weight.device),
                                 persistent=True)
            self.int_zero_point_in_comp[...] = -(zero_point_in.to(torch.float) * self.weight.to(torch.float)).sum(2) \
                .round().to(torch.int32)
        self.bias[:] = (conv.bias / (scale_in * scale_weight)).round().to(torch.int32)
        self._import_slope(prelu)
        if self.out_scaled_int:
            assert zero_point_out.dtype in (torch.int32, torch.int64) and zero_point_out.numel() == 1
            _fill_requant(self, scale_in * scale_weight / scale_out, zero_point_out)
        else:
            assert zero_point_out is None
            _fill_requant(self, scale_in * scale_weight, None)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        if prefix + 'int_zero_point_in_comp' in state_dict:
            self.use_zero_point_in = True
            self.register_buffer('int_zero_point_in_comp',
                                 torch.zeros((self.kernel_volume, self.out_ch), dtype=torch.int32), persistent=True)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def unique(self, *args, **kwargs):
        return torch.unique(*args, **kwargs)

    def forward(self, *args, **kwargs):
        if isinstance(args[0] if args else kwargs.get('input'), SparseTensor):
            return self.forward_with_sparse_tensor(*args, **kwargs)
        return self.forward_with_coords(*args, **kwargs)

    def forward_with_sparse_tensor(self, input: SparseTensor, residual=None, post_slope=None, post_requant=None, aux_out=None) -> SparseTensor:
        """`aux_out` (int8 [n_out, out_ch], with `post_requant`): dual output -- the int32 rows are returned as usual and the
        second stage's int8 rows are written to aux_out as well (ops.make_epilogue)."""
        caches = input._caches
        tag = (input.stride, self.kernel_size, self.stride)
        cur_kmap: Dict[str, Any] = caches.kmaps.get(tag)
        in_out_maps = cur_kmap.get('in_out_maps') if cur_kmap is not None else None
        hashmap_kv = caches.hashmaps.get(input.stride)
        if self.stride == (1, 1, 1):
            output_stride, output_coords, same = input.stride, input.C, True
        else:
            same = False
            output_stride = tuple(a * b for a, b in zip(input.stride, self.stride))
            if output_stride in caches.cmaps:
                output_coords = caches.cmaps[output_stride][0]
            elif (self.stride[0] & (self.stride[0] - 1)) == 0 and all(self.stride[0] == s for s in self.stride[1:]):
                output_coords = input.C.clone()
                output_coords[:, 1:] >>= (self.stride[0].bit_length() - 1)
                output_coords = self.unique(output_coords, dim=0)
            else:
                raise NotImplementedError((input.stride, self.stride))
        out_feats, hashmap_kv, in_out_maps = self.forward_with_coords(
            input.F, input.C, output_coords, in_out_maps, hashmap_kv, same, residual=residual, post_slope=post_slope,
            post_requant=post_requant, aux_out=aux_out)
        caches.kmaps.setdefault(tag, {}).setdefault('in_out_maps', in_out_maps)
        if hashmap_kv is not None:
            caches.hashmaps.setdefault(input.stride, hashmap_kv)
        caches.cmaps.setdefault(input.stride, (input.C, input.spatial_range))
        caches.cmaps.setdefault(output_stride, (output_coords, None))
        ret = SparseTensor(out_feats, output_coords, output_stride, None)
        ret._caches = caches
        return ret

    def forward_with_coords(self, in_feats, in_coords, out_coords, in_out_maps=None, hashmap_kv=None,
                            if_in_coords_equals_out_coords: bool = False, residual=None, post_slope=None, post_requant=None,
                            aux_out=None):
        """-> (out N2 x C2 int8 | Q8.23 int32, hashmap_kv, in_out_maps); conv + epilogue in one kernel."""
        ep = self.epilogue(True, residual, post_slope, post_requant=post_requant, aux_out=aux_out)
        if self.in_ch < 32 and not self.use_zero_point_in:
            # thin input (first conv C_in = 1, occupancy embeds C_in = 8): im2col + one tensor-core linear
            kv = self.kernel_volume
            if in_out_maps is None:
                in_out_maps, hashmap_kv = build_kernel_map(in_coords, out_coords, self.kernel_size, self.stride, hashmap_kv)
            kmap = _as_kernel_map(in_out_maps, out_coords.shape[0], kv, -1, in_feats.device)
            kp = max(32, (kv * self.in_ch + 15) // 16 * 16)
            def make():
                wf = torch.zeros((self.out_ch, kp), dtype=torch.int8, device=self.weight.device)
                wf[:, :kv * self.in_ch] = self.weight.permute(1, 0, 2).reshape(self.out_ch, kv * self.in_ch)
                return (wf,)

            cache = _cached(self, '_patch_weight', (self.weight._version, self.weight.data_ptr(), kp), make)
            patches = ops.gather_patches(in_feats, kmap.table, kp)
            return ops.linear(patches, cache[1], ep), hashmap_kv, kmap
        return sparse_conv_in8w8out32(
            in_feats, self.weight, in_coords, out_coords, self.kernel_size, self.stride, in_out_maps, hashmap_kv,
            self.int_zero_point_in_comp if self.use_zero_point_in else None, if_in_coords_equals_out_coords, _epilogue=ep)


class SparseConvIn8W8Out8(SparseConvIn8Out8):
    def __init__(self, in_ch, out_ch, kernel_size=(3, 3, 3), stride=(1, 1, 1), *args, **kwargs):
        super().__init__(in_ch, out_ch, kernel_size, stride, False, True, *args, **kwargs)

    @torch.no_grad()
    def import_parameters(self, scale_in, zero_point_in, scale_out, zero_point_out, conv):
        super().import_parameters(scale_in, zero_point_in, scale_out, zero_point_out, conv, None)


class SparseConvIn8W8Out32(SparseConvIn8Out8):
    def __init__(self, in_ch, out_ch, kernel_size=(3, 3, 3), stride=(1, 1, 1), *args, **kwargs):
        super().__init__(in_ch, out_ch, kernel_size, stride, False, False, *args, **kwargs)

    @torch.no_grad()
    def import_parameters(self, scale_in, zero_point_in, conv):
        super().import_parameters(scale_in, zero_point_in, None, None, conv, None)


class SparseConvPReLUIn8W8Out8(SparseConvIn8Out8):
    def __init__(self, in_ch, out_ch, kernel_size=(3, 3, 3), stride=(1, 1, 1), *args, **kwargs):
        super().__init__(in_ch, out_ch, kernel_size, stride, True, True, *args, **kwargs)

    @torch.no_grad()
    def import_parameters(self, scale_in, zero_point_in, scale_out, zero_point_out, conv, prelu):
        super().import_parameters(scale_in, zero_point_in, scale_out, zero_point_out, conv, prelu)


class SparseConvPReLUIn8W8Out32(SparseConvIn8Out8):
    def __init__(self, in_ch, out_ch, kernel_size=(3, 3, 3), stride=(1, 1, 1), *args, **kwargs):
        super().__init__(in_ch, out_ch, kernel_size, stride, True, False, *args, **kwargs)

    @torch.no_grad()
    def import_parameters(self, scale_in, zero_point_in, conv, prelu):
        super().import_parameters(scale_in, zero_point_in, None, None, conv, prelu)


class PReLUIn32Out32(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer('slope', torch.zeros((1,), dtype=torch.int32), persistent=True)

    @torch.no_grad()
    def import_parameters(self, prelu: nn.PReLU):
        assert prelu.weight.dtype == torch.float32 and prelu.weight.numel() == 1
        self.slope[:] = (prelu.weight * (1 << 25)).round().to(torch.int32)  # Q6.25

    def forward(self, input: torch.Tensor):
        return ops.prelu_i32(input, self.slope)

    def slope_in_unit_range(self) -> bool:
        """0 <= slope <= 1.0 (Q6.25), read back once per parameter version"""
        key = self.slope._version
        cache = getattr(self, '_unit_range', None)
        if cache is None or cache[0] != key:
            v = int(self.slope.item())
            cache = (key, 0 <= v <= (1 << 25))
            self._unit_range = cache
        return cache[1]


class RequantFxpToScaledInt8(LoadSaveUint32RequantMul):
    def __init__(self, eps=None, requant_mul_guard_bits=None):
        super().__init__()
        self.register_buffer('requant_mul', torch.zeros((1,), dtype=torch.uint32), persistent=True)
        self.register_buffer('requant_shift', torch.zeros((1,), dtype=torch.int32), persistent=True)
        self.register_buffer('int_zero_point_out', torch.zeros((1,), dtype=torch.int64), persistent=True)
        self.register_buffer('scale_out', torch.zeros((1,), dtype=torch.float32) - 1, persistent=True)
        self.register_buffer('zero_point_out', torch.zeros((1,), dtype=torch.float32), persistent=True)
        self.eps = eps if eps is not None else torch.finfo(torch.float32).eps
        self.requant_mul_guard_bits = requant_mul_guard_bits if requant_mul_guard_bits is not None else 2

    @torch.no_grad()
    def import_parameters(self, scale_out: torch.Tensor, zero_point_out: torch.Tensor):
        assert scale_out.dtype == torch.float32 and scale_out.numel() == 1
        if scale_out <= self.eps:
            print(f'Warning: {scale_out}')
            scale_out = scale_out.clip(min=self.eps)
        self.scale_out[:] = scale_out
        assert zero_point_out.dtype in (torch.int32, torch.int64) and zero_point_out.numel() == 1
        self.requant_shift[:] = torch.log2((1 << (32 - self.requant_mul_guard_bits)) * scale_out).floor()
        assert self.requant_shift >= 0, self.requant_shift
        self.requant_mul[:] = ((2 ** self.requant_shift.to(torch.float)) / scale_out).round().to(torch.int64).to(torch.uint32)
        self.zero_point_out[:] = zero_point_out
        self.int_zero_point_out[:] = (zero_point_out * (2 ** (SharedFxpShift + self.requant_shift.to(torch.float)))) \
            .round().to(torch.int64)

    def forward(self, input: torch.Tensor, prelu: Optional['PReLUIn32Out32'] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(Q8.23 x Qx.xx) >> (23 + xx) -> scaled int8; the single multiplier is broadcast by the kernel.
        `out`: an int8 [rows, ch] column slice of a wider buffer to write into (concatenation without a copy).
        `prelu`: a PReLUIn32Out32 that precedes this requant; its pass is folded into this kernel when its slope
        lies in [0, 1] (then the int32 saturation of the stand-alone PReLU cannot trigger, so the integers agree)."""
        slope = None
        if prelu is not None:
            if prelu.slope_in_unit_range():
                slope = prelu.slope
            else:
                input = prelu(input)
        ep = ops.make_epilogue(self.requant_mul, self.int_zero_point_out, SharedFxpShift + _shift_of(self), ops.OUT_I8, slope=slope)
        return ops.requant(input, ep, out=out)

    def as_post_stage(self, prelu: Optional['PReLUIn32Out32'] = None):
        """This requant (and a PReLUIn32Out32 in front of it) as the fused second stage of the producing int32 layer
        (ops.make_epilogue(post_requant=...)), or None when it cannot be fused with identical integers (slope outside
        [0, 1]: the stand-alone PReLU saturates to int32 first)."""
        if prelu is not None and not prelu.slope_in_unit_range():
            return None
        return (self.requant_mul, self.int_zero_point_out, SharedFxpShift + _shift_of(self), prelu.slope if prelu is not None else None)

    def bit_levels(self):
        """int8 images of the Q8.23 values 0 and 1.0 (an occupancy bit << 23) under this requant, cached."""
        key = (self.requant_mul._version, self.requant_shift._version, self.int_zero_point_out._version)
        cache = getattr(self, '_bit_levels', None)
        if cache is None or cache[0] != key:
            probe = torch.tensor([[0, 1 << SharedFxpShift, 0, 0]], dtype=torch.int32, device=self.requant_mul.device)
            q = self.forward(probe)[0, :2].tolist()
            cache = (key, int(q[0]), int(q[1]))
            self._bit_levels = cache
        return cache[1], cache[2]


class LinearIn8W8(_AffineIn8):
    def __init__(self, in_ch: int, out_ch: int, with_prelu: bool, out_scaled_int: bool, eps=None, requant_mul_guard_bits=None):
        super().__init__()
        self._register((out_ch, in_ch), out_ch, with_prelu)
        self.in_ch, self.out_ch = in_ch, out_ch
        self.with_prelu = with_prelu
        self.out_scaled_int = out_scaled_int
        self.eps = eps if eps is not None else torch.finfo(torch.float32).eps
        self.requant_mul_guard_bits = requant_mul_guard_bits if requant_mul_guard_bits is not None else 7

    @torch.no_grad()
    def import_parameters(self, scale_in, zero_point_in, scale_out, zero_point_out, linear: nn.Linear, prelu):
        """cuda_ops.py:541-607"""
        assert linear.weight.dtype == torch.float32 and linear.bias is not None
        assert zero_point_in.dtype in (torch.int32, torch.int64) and zero_point_in.numel() == 1
        scale_in, scale_out = self._import_scales(scale_in, scale_out)
        scale_weight = linear.weight.abs().amax(dim=1) / WeightRange
        if (scale_weight <= self.eps).any():
            print(f'Warning: {scale_weight}')
            scale_weight = scale_weight.clip(min=self.eps)
        self.scale_weight[:] = scale_weight
        self.weight[...] = (linear.weight / scale_weight[:, None]).round().clip(-WeightRange, WeightRange).to(torch.int8)
        self.zero_point_in[:] = zero_point_in
        self.bias[:] = (linear.bias / (scale_in * scale_weight)
                        - (zero_point_in.to(torch.float) * self.weight.to(torch.float)).sum(1)).round().to(torch.int32)
        self._import_slope(prelu)
        if self.out_scaled_int:
            assert zero_point_out.dtype in (torch.int32, torch.int64) and zero_point_out.numel() == 1
            _fill_requant(self, scale_in * scale_weight / scale_out, zero_point_out)
        else:
            assert zero_point_out is None
            _fill_requant(self, scale_in * scale_weight, None)

    def can_emit_aux(self) -> bool:
        """int32 output on the tensor-core kernel: its epilogue can also emit a consumer's int8 rows (dual output)"""
        return (not self.out_scaled_int and not getattr(self, 'padded_output', False) and self.out_ch % 16 == 0
                and ops.gemm_engine(self.in_ch if self.in_ch % 16 == 0 else self.in_ch - 8, self.out_ch) == 'tc')

    def forward(self, input: torch.Tensor, sel=None, n_out_rows=None, post_requant=None, aux_requant=None, out=None) -> torch.Tensor:
        """GEMM + bias + [PReLU] + requant in one kernel.  `sel` (from ops.slot_pairs) evaluates only the
        occupied (row, child) blocks of a C -> 8C linear: identical values, 4-8x less work.  `post_requant`
        (RequantFxpToScaledInt8.as_post_stage of the ONLY consumer of this layer's int32 output) makes the kernel
        emit that consumer's int8 directly: the Q8.23 tensor is never written."""
        if aux_requant is not None:
            # dual output: the Q8.23 rows AND `aux_requant` (a RequantFxpToScaledInt8.as_post_stage of one consumer) of them
            assert self.can_emit_aux() and post_requant is None and sel is None
            aux = torch.empty((input.shape[0], self.out_ch), dtype=torch.int8, device=input.device)
            return ops.linear(input, self.weight, self.epilogue(True, post_requant=aux_requant, aux_out=aux)), aux
        if post_requant is not None:
            assert not self.out_scaled_int and not getattr(self, 'padded_output', False)
            # `out`: an int8 column slice of a wider buffer the rows go to (the next Linear(cat(F, bits)), forward_cat_bits)
            return ops.linear(input, self.weight, self.epilogue(True, post_requant=post_requant), sel=sel, n_out_rows=n_out_rows, out=out)
        if sel is None and getattr(self, 'padded_output', False) and self.out_ch % 16 != 0 and self.in_ch % 16 == 0 and self.in_ch >= 32:
            # opt-in (the consumer must accept a row pitch), e.g. the 255 logits feeding the CDF kernels: run the
            # kernel on 256 zero-padded channels so that its rows are 16-byte aligned
            # vector stores, and hand back the [m, out_ch] column slice (row pitch 256) of that buffer.
            w, ep = self._padded_channels()
            return ops.linear(input, w, ep)[:, :self.out_ch]
        return ops.linear(input, self.weight, self.epilogue(True), sel=sel, n_out_rows=n_out_rows)

    def _padded_channels(self):
        key = (self.weight._version, self.bias._version, self.requant_mul._version, self.int_zero_point_out._version,
               self.weight.data_ptr())
        def make():
            n_pad = (self.out_ch + 15) // 16 * 16
            pad = n_pad - self.out_ch
            w = torch.nn.functional.pad(self.weight, (0, 0, 0, pad)).contiguous()
            bias = torch.nn.functional.pad(self.bias, (0, pad)).contiguous()
            mul = self.requant_mul
            if mul.numel() > 1:
                mul = torch.nn.functional.pad(mul.view(torch.int32), (0, pad)).contiguous().view(torch.uint32)
            shift = _shift_of(self)
            out_type = ops.OUT_I8 if self.out_scaled_int else ops.OUT_I32
            if not self.out_scaled_int:
                shift -= SharedFxpShift
            ep = ops.make_epilogue(mul, self.int_zero_point_out, shift, out_type, bias=bias,
                                   slope=self.slope if self.with_prelu else None)
            return w, ep

        cache = _cached(self, '_pad_cache', key, make)
        return cache[1], cache[2]

    def can_cat_bits(self) -> bool:
        """Linear over cat(F, 8 occupancy bits) as ONE tensor-core GEMM with K = C + 16 (forward_cat_bits)"""
        c = self.in_ch - 8
        return c >= 32 and c % 16 == 0 and ops.gemm_engine(c + 16, self.out_ch) == 'tc'

    def forward_cat_bits(self, buf: torch.Tensor) -> torch.Tensor:
        """Linear over cat(input, bits) with the concatenation already in memory: `buf` is int8 [rows, C + 16] whose first C
        columns hold the requantised features (written there by their producer: ops.requant(out=...) or a fused second stage
        with an output pitch), columns C .. C+7 the requantised occupancy bits (ops.occ_bits_q8) and 8 zero columns.  The
        occupancy row-bias form (forward_with_bits) keeps K = C but reads 1 KB of bias per output row in the epilogue -- as
        much as an int32 residual; here the bit channels ride in a third, nearly empty K stage of the MMA instead."""
        c = self.in_ch - 8
        key = (self.weight._version, self.weight.data_ptr())

        def make():
            return (torch.nn.functional.pad(self.weight, (0, 8)).contiguous(),)   # [out_ch, C + 8 + 8]

        _, w_cat = _cached(self, '_cat_cache', key, make)
        assert buf.dtype == torch.int8 and buf.shape[1] == c + 16 and buf.is_contiguous()
        return ops.linear(buf, w_cat, self.epilogue(True))

    def forward_with_bits(self, input: torch.Tensor, occ: torch.Tensor, q0: int, q1: int, aux_requant=None) -> torch.Tensor:
        """Linear over cat(input, bits) where `bits` are the 8 occupancy channels of `occ` (channel k = bit 7-k)
        already requantised to the two int8 values q0 (bit clear) / q1 (bit set).  Their contribution
        sum_k q[bit_k] * W[:, C+k] depends only on the occupancy byte, so it is a 256-row bias table and the
        contraction keeps K = C (identical integers, no concat, tensor-core friendly K)."""
        C = self.in_ch - 8
        key = (q0, q1, self.weight._version, self.weight.data_ptr())
        def make():
            w = self.weight
            pat = ((torch.arange(256, device=w.device)[:, None] >> torch.arange(7, -1, -1, device=w.device)[None]) & 1).double()
            q = pat * float(q1) + (1.0 - pat) * float(q0)                       # [256, 8]
            table = (q @ w[:, C:].double().T).round().to(torch.int32).contiguous()  # exact: |values| < 2^53
            return w[:, :C].contiguous(), table, int(table.abs().max().item())

        _, w_main, table, bound = _cached(self, '_bits_cache', key, make)
        if aux_requant is not None:  # dual output, see forward()
            aux = torch.empty((input.shape[0], self.out_ch), dtype=torch.int8, device=input.device)
            return ops.linear(input, w_main, self.epilogue(True, row_bias=(table, occ, bound), post_requant=aux_requant, aux_out=aux)), aux
        return ops.linear(input, w_main, self.epilogue(True, row_bias=(table, occ, bound)))


class LinearIn8W8Out8(LinearIn8W8):
    def __init__(self, in_ch, out_ch, *args, **kwargs):
        super().__init__(in_ch, out_ch, False, True, *args, **kwargs)

    @torch.no_grad()
    def import_parameters(self, scale_in, zero_point_in, scale_out, zero_point_out, linear):
        super().import_parameters(scale_in, zero_point_in, scale_out, zero_point_out, linear, None)


class LinearIn8W8Out32(LinearIn8W8):
    def __init__(self, in_ch, out_ch, *args, **kwargs):
        super().__init__(in_ch, out_ch, False, False, *args, **kwargs)

    @torch.no_grad()
    def import_parameters(self, scale_in, zero_point_in, linear):
        super().import_parameters(scale_in, zero_point_in, None, None, linear, None)


class LinearPReLUIn8W8Out8(LinearIn8W8):
    def __init__(self, in_ch, out_ch, *args, **kwargs):
        super().__init__(in_ch, out_ch, True, True, *args, **kwargs)

    @torch.no_grad()
    def import_parameters(self, scale_in, zero_point_in, scale_out, zero_point_out, linear, prelu):
        super().import_parameters(scale_in, zero_point_in, scale_out, zero_point_out, linear, prelu)


class LinearPReLUIn8W8Out32(LinearIn8W8):
    def __init__(self, in_ch, out_ch, *args, **kwargs):
        super().__init__(in_ch, out_ch, True, False, *args, **kwargs)

    @torch.no_grad()
    def import_parameters(self, scale_in, zero_point_in, linear, prelu):
        super().import_parameters(scale_in, zero_point_in, None, None, linear, prelu)


class SparseResBlockIn32W8Out32(nn.Module):
    def __init__(self, ch: int, eps=None):
        super().__init__()
        self.ch = ch
        self.eps = eps if eps is not None else torch.finfo(torch.float32).eps
        self.input_requant = RequantFxpToScaledInt8()
        self.conv_prelu = SparseConvPReLUIn8W8Out8(ch, ch, (3, 3, 3), (1, 1, 1))
        self.conv2 = SparseConvIn8W8Out32(ch, ch, (3, 3, 3), (1, 1, 1))
        self.prelu = PReLUIn32Out32()

    @torch.no_grad()
    def import_parameters(self, block):
        scale, zero_point = block.obs.calculate_qparams()
        scale2, zero_point2 = block.obs2.calculate_qparams()
        self.input_requant.import_parameters(scale, zero_point)
        self.conv_prelu.import_parameters(scale, zero_point, scale2, zero_point2, block.conv, block.act)
        self.conv2.import_parameters(scale2, zero_point2, block.conv2)
        self.prelu.import_parameters(block.act2)

    def can_fuse_consumer(self) -> bool:
        """conv2 runs on the tensor-core kernel (the only one with the fused second stage)"""
        return not self.conv2.use_zero_point_in and ops.gemm_engine(self.ch, self.ch, 27) == 'tc'

    def forward(self, input: SparseTensor, post_requant=None, input_q: Optional[torch.Tensor] = None, aux_requant=None):
        """cuda_ops.py:82-92.  The residual add (int32 wrap) and the final PReLU ride in conv2's epilogue.
        `post_requant` (RequantFxpToScaledInt8.as_post_stage of the block's ONLY consumer): the block then returns
        that consumer's int8 rows and its Q8.23 output is never written (check can_fuse_consumer first)."""
        # `input_q`: input_requant(input.F) already emitted by the producer of input.F (its dual output);
        # `aux_requant`: ONE further consumer's requant of the block output, emitted by conv2 beside the Q8.23 rows
        # (returns (block output, int8 rows) then)
        x = SparseTensor(self.input_requant(input.F) if input_q is None else input_q, input.C, input.stride, input.spatial_range)
        x._caches = input._caches
        x = self.conv_prelu(x)
        aux = None
        if aux_requant is not None:
            assert post_requant is None and self.can_fuse_consumer()
            aux = torch.empty((input.F.shape[0], self.ch), dtype=torch.int8, device=input.F.device)
            x = self.conv2.forward_with_sparse_tensor(x, residual=input.F, post_slope=self.prelu.slope, post_requant=aux_requant, aux_out=aux)
        else:
            x = self.conv2.forward_with_sparse_tensor(x, residual=input.F, post_slope=self.prelu.slope, post_requant=post_requant)
        assert input.F.dtype == torch.int32 and x.F.dtype == (torch.int8 if post_requant is not None else torch.int32)
        out = SparseTensor(x.F, input.C, input.stride, input.spatial_range)
        out._caches = input._caches
        return out if aux_requant is None else (out, aux)


def softmax_int32(input: torch.Tensor) -> torch.Tensor:
    return ops.softmax_i32(input)
