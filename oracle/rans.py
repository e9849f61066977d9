"""ctypes front-end of oracle/rans_oracle.c with the reference's Python class surface.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Mirrors (same method names / argument meaning):
  RansEncoder, RansDecoder ........ models/convolutional/lossy_coord_v3/rans_coder/simple_rans_wrapper.cpp:272-286
  IndexedRansCoder, BinaryRansCoder,
  batched_pmf_to_quantized_cdf .... lib/entropy_models/rans_coder/rans_wrapper.cpp:430-451
"""
import ctypes as C
import os.path as osp
import subprocess

import numpy as np

_HERE = osp.dirname(osp.abspath(__file__))
_SO = osp.join(_HERE, '_build', 'librans_oracle.so')
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not osp.isfile(_SO) or osp.getmtime(_SO) < osp.getmtime(osp.join(_HERE, 'rans_oracle.c')):
            subprocess.run(['make', '-C', _HERE], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(_SO)
        vp, sz, u64, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int32
        L.fo_simple_enc_new.restype = vp; L.fo_simple_enc_new.argtypes = [sz]
        L.fo_simple_enc_free.argtypes = [vp]
        L.fo_simple_enc_encode.restype = u64
        L.fo_simple_enc_encode.argtypes = [vp, vp, sz, sz, vp, sz]
        L.fo_simple_enc_encode_bin.restype = u64
        L.fo_simple_enc_encode_bin.argtypes = [vp, vp, sz, vp, sz]
        L.fo_simple_enc_flush.restype = sz; L.fo_simple_enc_flush.argtypes = [vp, vp, sz]
        L.fo_simple_dec_new.restype = vp
        L.fo_simple_dec_free.argtypes = [vp]
        L.fo_simple_dec_flush.argtypes = [vp, vp]
        L.fo_simple_dec_decode.argtypes = [vp, vp, sz, sz, vp, sz]
        L.fo_simple_dec_decode_bin.argtypes = [vp, vp, sz, vp, sz]
        L.fo_indexed_encode.restype = sz
        L.fo_indexed_encode.argtypes = [vp, vp, vp, sz, vp, C.c_int, vp, vp, sz, vp, sz]
        L.fo_indexed_decode.argtypes = [vp, vp, vp, sz, vp, C.c_int, vp, vp, vp, sz]
        L.fo_binary_encode.restype = sz; L.fo_binary_encode.argtypes = [vp, vp, sz, vp, sz]
        L.fo_binary_decode.argtypes = [vp, vp, vp, sz]
        L.fo_pmf_to_quantized_cdf.restype = i32
        L.fo_pmf_to_quantized_cdf.argtypes = [vp, sz, vp, C.c_int, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u16(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint16, a.dtype
    return a


class RansEncoder:
    def __init__(self, enc_buf_size=32 * 1024 * 1024):
        self._cap = int(enc_buf_size)
        self._h = lib().fo_simple_enc_new(self._cap)

    def __del__(self):
        if getattr(self, '_h', None):
            lib().fo_simple_enc_free(self._h)
            self._h = None

    def encode(self, cdf_arr, symbol_arr):
        cdf, sym = _u16(cdf_arr), _u16(symbol_arr)
        assert cdf.ndim == 2 and sym.ndim == 1
        assert cdf.shape[0] in (1, sym.shape[0])
        return int(lib().fo_simple_enc_encode(self._h, _p(cdf), cdf.shape[0], cdf.shape[1], _p(sym), sym.shape[0]))

    def encode_bin(self, cdf_arr, symbol_arr):
        cdf = _u16(cdf_arr).reshape(-1)
        sym = np.ascontiguousarray(symbol_arr, dtype=np.bool_).view(np.uint8)
        assert cdf.shape[0] in (1, sym.shape[0])
        return int(lib().fo_simple_enc_encode_bin(self._h, _p(cdf), cdf.shape[0], _p(sym), sym.shape[0]))

    def flush(self):
        out = np.empty(self._cap, dtype=np.uint8)
        n = lib().fo_simple_enc_flush(self._h, _p(out), self._cap)
        return out[:n].tobytes()


class RansDecoder:
    def __init__(self):
        self._h = lib().fo_simple_dec_new()
        self._keep = None

    def __del__(self):
        if getattr(self, '_h', None):
            lib().fo_simple_dec_free(self._h)
            self._h = None

    def flush(self, encoded: bytes):
        self._keep = np.frombuffer(bytes(encoded) + b'\0' * 8, dtype=np.uint8)  # padded private copy
        lib().fo_simple_dec_flush(self._h, _p(self._keep))
        return 0

    def decode(self, cdf_arr, symbol_arr):
        cdf = _u16(cdf_arr)
        assert symbol_arr.dtype == np.uint16 and symbol_arr.flags.c_contiguous and symbol_arr.flags.writeable
        assert cdf.shape[0] in (1, symbol_arr.shape[0])
        lib().fo_simple_dec_decode(self._h, _p(cdf), cdf.shape[0], cdf.shape[1], _p(symbol_arr), symbol_arr.shape[0])
        return 0

    def decode_bin(self, cdf_arr, symbol_arr):
        cdf = _u16(cdf_arr).reshape(-1)
        assert symbol_arr.dtype == np.bool_ and symbol_arr.flags.c_contiguous
        lib().fo_simple_dec_decode_bin(self._h, _p(cdf), cdf.shape[0], _p(symbol_arr.view(np.uint8)), symbol_arr.shape[0])
        return 0


def batched_pmf_to_quantized_cdf(pmf_array, offset_array, overflow_coding):
    """cdf_ops.cpp:111-143.  Mutates `offset_array` in place in overflow mode (and `pmf_array`
    becomes its prefix sum), exactly like the reference."""
    assert pmf_array.dtype == np.float64 and pmf_array.ndim == 2 and pmf_array.flags.c_contiguous
    assert offset_array.dtype == np.int32 and offset_array.ndim == 1
    cdfs = []
    tmp = np.empty(pmf_array.shape[1] + 2, dtype=np.uint32)
    for i in range(pmf_array.shape[0]):
        row = pmf_array[i]
        n = lib().fo_pmf_to_quantized_cdf(_p(row), row.shape[0], _p(offset_array[i:i + 1]), int(bool(overflow_coding)), _p(tmp))
        assert n > 0, 'no symbol to steal frequency from'
        cdfs.append(tmp[:n].tolist())
    return cdfs


class IndexedRansCoder:
    def __init__(self, overflow_coding: bool, batch_size: int, enc_buf_size=8 * 1024 * 1024):
        assert batch_size > 0
        self.overflow_coding = bool(overflow_coding)
        self.batch_size = int(batch_size)
        self.cdfs = None
        self.offset_array = None

    def init_with_pmfs(self, pmf_array, offset_array):
        cdfs = batched_pmf_to_quantized_cdf(pmf_array, offset_array, self.overflow_coding)
        return self.init_with_quantized_cdfs(cdfs, offset_array)

    def init_with_quantized_cdfs(self, cdfs, offset_array):
        self.cdfs = [list(map(int, c)) for c in cdfs]
        self.offset_array = np.ascontiguousarray(offset_array, dtype=np.int32)
        self._flat = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.uint32) for c in self.cdfs]))
        lens = np.array([len(c) for c in self.cdfs], dtype=np.int32)
        self._len = lens
        self._off = np.concatenate([[0], np.cumsum(lens[:-1])]).astype(np.int64)
        return 0

    def get_cdfs(self):
        return self.cdfs

    def get_offset_array(self):
        return self.offset_array

    def _enc(self, sym, idx):
        sym = np.ascontiguousarray(sym, dtype=np.int32)
        assert sym.ndim == 2 and sym.shape[0] == self.batch_size
        if idx is not None:
            idx = np.ascontiguousarray(idx, dtype=np.int32)
            assert idx.shape == sym.shape
        out = []
        n = sym.shape[1]
        cap = 4 * n * (17 if self.overflow_coding else 1) + 4096
        buf = np.empty(cap, dtype=np.uint8)
        for b in range(self.batch_size):
            start = lib().fo_indexed_encode(
                _p(self._flat), _p(self._off), _p(self._len), len(self.cdfs), _p(self.offset_array),
                int(self.overflow_coding), _p(sym[b]), _p(idx[b]) if idx is not None else None, n, _p(buf), cap)
            assert start != C.c_size_t(-1).value, 'oracle encode buffer too small'
            out.append(buf[start:].tobytes())
        return out

    def encode(self, symbol_array):
        return self._enc(symbol_array, None)

    def encode_with_indexes(self, symbol_array, index_array):
        return self._enc(symbol_array, index_array)

    def _dec(self, encoded_list, idx, out):
        assert out.dtype == np.int32 and out.ndim == 2 and out.flags.c_contiguous
        assert len(encoded_list) == self.batch_size
        if idx is not None:
            idx = np.ascontiguousarray(idx, dtype=np.int32)
        for b in range(self.batch_size):
            data = np.frombuffer(bytes(encoded_list[b]) + b'\0' * 8, dtype=np.uint8)
            lib().fo_indexed_decode(
                _p(self._flat), _p(self._off), _p(self._len), len(self.cdfs), _p(self.offset_array),
                int(self.overflow_coding), _p(data), _p(idx[b]) if idx is not None else None, _p(out[b]), out.shape[1])
        return 0

    def decode(self, encoded_list, symbol_array):
        return self._dec(encoded_list, None, symbol_array)

    def decode_with_indexes(self, encoded_list, index_array, symbol_array):
        return self._dec(encoded_list, index_array, symbol_array)


class BinaryRansCoder:
    def __init__(self, batch_size: int, enc_buf_size=8 * 1024 * 1024):
        assert batch_size > 0
        self.batch_size = int(batch_size)

    def encode(self, symbol_array, prob_array):
        sym = np.ascontiguousarray(symbol_array, dtype=np.bool_).view(np.uint8)
        prob = np.ascontiguousarray(prob_array, dtype=np.uint32)
        assert sym.ndim == 2 and sym.shape == prob.shape and sym.shape[0] == self.batch_size
        n = sym.shape[1]
        cap = 4 * n + 64
        buf = np.empty(cap, dtype=np.uint8)
        out = []
        for b in range(self.batch_size):
            start = lib().fo_binary_encode(_p(sym[b]), _p(prob[b]), n, _p(buf), cap)
            assert start != C.c_size_t(-1).value
            out.append(buf[start:].tobytes())
        return out

    def decode(self, encoded_list, prob_array, symbol_array):
        prob = np.ascontiguousarray(prob_array, dtype=np.uint32)
        assert symbol_array.dtype == np.bool_ and symbol_array.shape == prob.shape
        for b in range(self.batch_size):
            data = np.frombuffer(bytes(encoded_list[b]) + b'\0' * 8, dtype=np.uint8)
            lib().fo_binary_decode(_p(data), _p(prob[b]), _p(symbol_array[b].view(np.uint8)), prob.shape[1])
        return 0
