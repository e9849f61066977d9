"""Drop-in for the reference's JIT-built module `space_filling_curves_ext`
(lib/space_filling_curves/src/morton3d.cu:39-76, loaded by lib/space_filling_curves/__init__.py:16-40): the one entry
point the codec path calls, `morton3d_encode_magicbits(coords, axis_order)`, forwarded to fpcc_morton_encode."""
import torch

from . import ops

# axis_order names the coordinate that lands in the LEAST significant interleaved bit first (morton3d.cu:30-36);
# the codec calls it with inverse=True -> 'zyx' (x most significant; lossl_coord_int/model.py:398)
_MSB_AXIS = {'zyx': 0, 'xyz': 2}


def morton3d_encode_magicbits(coords: torch.Tensor, axis_order: str) -> torch.Tensor:
    if not coords.is_cuda:
        raise RuntimeError('Expected coords.is_cuda()')
    if coords.dim() != 2 or coords.size(1) != 3:
        raise RuntimeError('Expected coords.dim() == 2 && coords.size(1) == 3')
    if coords.dtype != torch.int32:
        raise RuntimeError('Expected coords.scalar_type() == at::kInt')
    if axis_order not in _MSB_AXIS:
        raise RuntimeError(f'axis order {axis_order!r}: only xyz / zyx are on the codec path')
    n = coords.size(0)
    if n and (coords.stride(1) != 1 or coords.stride(0) < 3):
        coords = coords.contiguous()
    codes = torch.empty(n, dtype=torch.int64, device=coords.device)
    if n:
        ops._call('fpcc_morton_encode', coords.data_ptr(), coords.stride(0), n, _MSB_AXIS[axis_order], codes.data_ptr(), ops._s())
    return codes
