"""CPU MinkowskiEngine stand-in for the float oracle -- TEST INFRASTRUCTURE ONLY.

MinkowskiEngine cannot be installed offline, so the float codecs of the reference (lossy_coord_v2, ...) have no
runnable backend in this container.  This module is the shim of fastpcc_b200/me.py (ME's public surface as the
reference uses it: coordinate manager, sparse tensor, layer classes) with its operator front-end `ops` rebound to
oracle/float_ops_cpu.py (torch-CPU fp32 restatements of the convolution / linear / kernel-map semantics) and the compute
dtype set to float32.  The reference's unmodified model code runs on it here, which gives the float path an fp32
result to compare the fp16 tensor-core kernels against (tests/golden/make_lossy_golden.py, tests/test_gpu_lossy_dropin.py).

What is independent of the product and what is not: every ARITHMETIC operator is restated here; the Python coordinate
bookkeeping (which key a strided / pruned / generated coordinate set gets, Morton ordering of new sets) is the shim's
own source, shared on purpose -- both backends must hand the reference code the same row order.  ME itself being
absent, parity of that bookkeeping with ME is unpinned.
"""
import os.path as osp
import sys
import types

import torch

from . import float_ops_cpu

_SRC = osp.join(osp.dirname(osp.dirname(osp.abspath(__file__))), 'fastpcc_b200', 'me.py')


def load(name='oracle_me_cpu'):
    """-> a fresh module object with ME's surface, CPU operators, float32 compute"""
    with open(_SRC) as f:
        src = f.read()
    src = src.replace('from . import ops\n', '')
    mod = types.ModuleType(name)
    mod.__file__ = _SRC
    mod.ops = float_ops_cpu
    exec(compile(src, _SRC, 'exec'), mod.__dict__)
    mod._COMPUTE = torch.float32
    mod.GROUP_ROWS_MIN = 1 << 62
    sub = types.ModuleType(name + '.MinkowskiSparseTensor')
    sub.SparseTensorQuantizationMode = mod.SparseTensorQuantizationMode
    sub.SparseTensor = mod.SparseTensor
    mod.MinkowskiSparseTensor = sub
    sys.modules[name] = mod
    sys.modules[name + '.MinkowskiSparseTensor'] = sub
    return mod
