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
