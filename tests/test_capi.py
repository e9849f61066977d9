"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a without a GPU, loads, and exports every
symbol that include/fastpcc_b200.h declares; the ctypes table covers exactly those symbols; entry points fail
loudly (no CPU fallback) when no device is present."""
import ctypes
import os.path as osp
import re

import pytest

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))


def _declared():
    text = open(osp.join(ROOT, 'include', 'fastpcc_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(fpcc_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from fastpcc_b200 import _lib, build
    so = build.build()
    assert osp.isfile(so)
    lib = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f'{n} declared in the header but not exported'
    assert sorted(_lib.SIGNATURES) == names, set(_lib.SIGNATURES) ^ set(names)
    _lib.load(build_if_missing=False)
    assert _lib.load().fpcc_version() >= 100


def test_every_entry_point_cites_the_reference_interface_it_replaces():
    text = open(osp.join(ROOT, 'include', 'fastpcc_b200.h')).read()
    for ref in ('hashmap_cuda.cuh', 'gather_gemm_scatter.cu', 'gemm.cu', 'softmax.cu', 'simple_rans_wrapper.cpp',
                'rans_wrapper.cpp', 'cdf_ops.cpp', 'cuda_ops.py', 'model.py', 'morton3d.cu'):
        assert ref in text, ref


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from fastpcc_b200 import _lib
    lib = _lib.load()
    rc = lib.fpcc_device_check(None, None, None)
    assert rc != 0 and lib.fpcc_last_error()
    from fastpcc_b200 import ops
    with pytest.raises(Exception):
        ops.softmax_i32(torch.zeros((2, 4), dtype=torch.int32))  # CPU tensors are rejected, nothing is computed on the host


def test_product_does_not_import_the_oracle():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import fastpcc_b200, fastpcc_b200.ops, fastpcc_b200.rans_coder, "
            "fastpcc_b200.lossl_coord_int, fastpcc_b200.int_sparse_conv.cuda_ops; "
            "bad=[m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; assert not bad, bad" % ROOT)
    subprocess.run([sys.executable, '-c', code], check=True)
    for dirpath, _, files in __import__('os').walk(osp.join(ROOT, 'fastpcc_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(osp.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_argument_checks_of_the_round2_entry_points_need_no_device():
    """Entry points added in round 2 validate their arguments on the host before any CUDA work: bad calls come back as an
    error code + message (the RuntimeError of the Python layer), with or without a GPU."""
    import ctypes as C
    from fastpcc_b200 import _lib
    lib = _lib.load()
    buf = (C.c_uint8 * 64)()
    p = C.addressof(buf)
    # tensor-core weight gradient: padded channels outside {128, 256} x multiple of 64
    assert lib.fpcc_spconv_wgrad_f16(p, p, 0, 96, 128, p, p, p, 1, p, 96, 128, None) != 0
    assert b'padded channels' in lib.fpcc_last_error()
    assert lib.fpcc_spconv_wgrad_f16(p, p, 2, 128, 128, p, p, p, 1, p, 128, 128, None) != 0 and b'dtype' in lib.fpcc_last_error()
    # occupancy-bit writer: pitch not a multiple of 16, levels outside int8
    assert lib.fpcc_occ_bits_q8(p, 4, 0, 1, p, 24, None) != 0 and b'16-byte' in lib.fpcc_last_error()
    assert lib.fpcc_occ_bits_q8(p, 4, 0, 300, p, 32, None) != 0 and b'int8' in lib.fpcc_last_error()
    # requant with an output pitch: one multiplier, no bias, int8 output, channels % 16 == 0
    ep = _lib.Epilogue()
    ep.requant_mul, ep.zero_point, ep.shift, ep.out_type, ep.mul_is_scalar = p, p, 40, 0, 1
    assert lib.fpcc_requant_ld(p, 4, 24, C.byref(ep), p, 32, None) != 0 and b'multiples of 16' in lib.fpcc_last_error()
    ep.mul_is_scalar = 0
    assert lib.fpcc_requant_ld(p, 4, 32, C.byref(ep), p, 32, None) != 0 and b'one multiplier' in lib.fpcc_last_error()
    # epilogue consistency: aux_out needs the second stage, out_ld applies to int8 rows only
    ep = _lib.Epilogue()
    ep.requant_mul, ep.zero_point, ep.shift, ep.out_type, ep.mul_is_scalar, ep.aux_out = p, p, 10, 2, 1, p
    assert lib.fpcc_linear_i8(p, 4, 32, p, 16, None, None, None, 1, 0, C.byref(ep), p, None) != 0 and b'aux_out' in lib.fpcc_last_error()
    ep.aux_out, ep.out_ld = None, 48
    assert lib.fpcc_linear_i8(p, 4, 32, p, 16, None, None, None, 1, 0, C.byref(ep), p, None) != 0 and b'out_ld' in lib.fpcc_last_error()
    # gather of coordinate rows: misaligned rows
    assert lib.fpcc_gather_rows16(p + 4, p, 1, p, None) != 0 and b'aligned' in lib.fpcc_last_error()
