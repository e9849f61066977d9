"""Byte containers of the codec path over the C ABI (fastpcc_b200/csrc/container.cu): frame header and partition
container of lossl_coord_int (model.py:447-463, 466-473, 510-521) and `BytesListUtils`
(lib/entropy_models/hyperprior/noisy_deep_factorized/utils.py:8-76) with the reference's class and method names."""
import ctypes as C
import io
from typing import List, Optional, Tuple

from . import _lib


def _err(what):
    msg = _lib.load().fpcc_last_error()
    raise RuntimeError(f'{what}: {msg.decode() if msg else "error"}')


def _items(bytes_list):
    n = len(bytes_list)
    ptrs = (C.c_char_p * n)(*bytes_list)
    lens = (C.c_int64 * n)(*[len(b) for b in bytes_list])
    return n, ptrs, lens


def write_frame_header(coord_offset, bottom_points: int) -> bytes:
    off = (C.c_int32 * 3)(*[int(v) for v in coord_offset])
    out = (C.c_uint8 * 8)()
    _lib.call('fpcc_frame_header_write', off, int(bottom_points), out)
    return bytes(out)


def read_frame_header(stream: bytes) -> Tuple[List[int], int, bytes]:
    """-> (coord_offset[3], bottom point count, rANS payload)"""
    off = (C.c_int32 * 3)()
    cnt = C.c_int()
    _lib.call('fpcc_frame_header_read', stream, len(stream), off, C.byref(cnt))
    return list(off), cnt.value, stream[8:]


def pack_partitions(streams: List[bytes]) -> bytes:
    n, ptrs, lens = _items(streams)
    lib = _lib.load()
    total = lib.fpcc_partitions_pack(ptrs, lens, n, None, 0)
    if total < 0:
        _err('pack_partitions')
    out = C.create_string_buffer(max(total, 1))
    if lib.fpcc_partitions_pack(ptrs, lens, n, out, total) != total:
        _err('pack_partitions')
    return out.raw[:total]


def split_partitions(blob: bytes) -> List[bytes]:
    lib = _lib.load()
    n = lib.fpcc_partitions_index(blob, len(blob), 0, None, None)
    if n < 0:
        _err('split_partitions')
    offs, lens = (C.c_int64 * max(n, 1))(), (C.c_int64 * max(n, 1))()
    if lib.fpcc_partitions_index(blob, len(blob), n, offs, lens) != n:
        _err('split_partitions')
    return [blob[offs[i]: offs[i] + lens[i]] for i in range(n)]


class BytesListUtils:
    @staticmethod
    def concat_bytes_list(bytes_list: List[bytes], bs_io: io.BytesIO = None) -> Optional[bytes]:
        assert len(bytes_list) > 1
        n, ptrs, lens = _items(bytes_list)
        lib = _lib.load()
        total = lib.fpcc_bytes_list_concat(ptrs, lens, n, None, 0)
        if total < 0:
            _err('concat_bytes_list')
        out = C.create_string_buffer(total)
        if lib.fpcc_bytes_list_concat(ptrs, lens, n, out, total) != total:
            _err('concat_bytes_list')
        if bs_io is None:
            return out.raw[:total]
        bs_io.write(out.raw[:total])

    @staticmethod
    def split_bytes_list(concat_bytes: Optional[bytes], bytes_list_len: int, bs_io: io.BytesIO = None) -> List[bytes]:
        if bs_io is not None:
            assert concat_bytes is None
            start = bs_io.tell()
            data = bs_io.read()
        else:
            data = concat_bytes
        offs, lens = (C.c_int64 * bytes_list_len)(), (C.c_int64 * bytes_list_len)()
        used = _lib.load().fpcc_bytes_list_split(data, len(data), bytes_list_len, offs, lens)
        if used < 0:
            _err('split_bytes_list')
        if bs_io is not None:
            bs_io.seek(start + used)  # the reference leaves the cursor right after the last item
        return [data[offs[i]: offs[i] + lens[i]] for i in range(bytes_list_len)]
