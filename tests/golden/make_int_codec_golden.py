"""Mints tests/golden/int_codec_golden.json by running the reference's OWN Python for the lossl_coord_int codec
(models/convolutional/lossl_coord_int/model.py + lib/int_sparse_conv/cuda_ops.py, imported unmodified from
/root/reference) on the CPU of this container.

What is the reference and what is not, in this run:
  * reference, unmodified: Model.compress / decompress / get_bin / batch_quantize_pmf_torch, OneScalePredictor,
    OneScaleMultiStepPredictor, every layer class of cuda_ops.py (kernel-map compaction, centre-offset omission,
    cache handling, requant shift arithmetic, state-dict loading), morton_encode_magicbits (its pure-torch
    branch), and the reference's compiled range coder (oracle/_ref/simple_rans_ext_cpp, built from the
    reference's C++ by oracle/build_ref.py).
  * stand-in: the 20 entry points of the CUDA extension `int_sparse_conv_ext` (binding.cu:114-145), which cannot
    be built or run here (needs a GPU, torchsparse and the author's CUTLASS fork).  They are emulated on the CPU by
    oracle/int_ops.py primitives (each restated from the .cu source it cites).  torchsparse.SparseTensor is a
    4-attribute stand-in (F, C, stride, _caches) -- the reference uses nothing else of it on this path.

So the fixture pins the whole-codec orchestration of oracle/lossl_coord_int.py and of the CUDA product to the
reference's Python byte for byte; the primitive kernels stay pinned to their sources (tests/test_oracle_int_ops.py).

Run:  python tests/golden/make_int_codec_golden.py     (needs /root/reference and oracle/_ref)
"""
import hashlib
import importlib
import json
import os.path as osp
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = osp.dirname(osp.abspath(__file__))
ROOT = osp.dirname(osp.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)

from oracle import build_ref, int_ops as K  # noqa: E402
from fastpcc_b200 import synth  # noqa: E402  (synthetic inputs / parameters only; no product code on the path)

from tests.golden.int_codec_cases import CASES, case_cloud  # noqa: E402


# ----------------------------------------------------------------------------------------------
# CPU stand-in for int_sparse_conv_ext (binding.cu:114-145)
# ----------------------------------------------------------------------------------------------

def _np(t):
    if t.dtype == torch.uint32:
        return t.view(torch.int32).numpy().view(np.uint32)
    return t.numpy()


def _make_ext():
    ext = types.ModuleType('int_sparse_conv_ext_cpu_standin')

    def cutlass_gemm_int8(A, B, C, D):
        D.copy_(torch.from_numpy(K.gemm_int8(_np(A), _np(B), _np(C) if C.numel() else None)))

    def cutlass_gather_gemm_scatter_int8(A, B, C, D, gather_idx, scatter_idx):
        assert C.data_ptr() == D.data_ptr()
        d = D.numpy()
        K.gather_gemm_scatter_int8(_np(A), _np(B), d, _np(gather_idx).astype(np.int64), _np(scatter_idx).astype(np.int64))

    def softmax_int32(x):
        out = K.softmax_int32(_np(x))
        return torch.from_numpy(out.view(np.int32)).view(torch.uint32)

    def _rq(dt):
        def plain(inp, mul, zp, shift):
            return torch.from_numpy(K.requant(_np(inp), _np(mul), _np(zp), shift, dt))

        def bias(inp, b, mul, zp, shift):
            return torch.from_numpy(K.requant(_np(inp), _np(mul), _np(zp), shift, dt, bias=_np(b)))

        def prelu(inp, sl, mul, zp, shift):
            return torch.from_numpy(K.requant(_np(inp), _np(mul), _np(zp), shift, dt, slope=_np(sl)))

        def bias_prelu(inp, b, sl, mul, zp, shift):
            return torch.from_numpy(K.requant(_np(inp), _np(mul), _np(zp), shift, dt, bias=_np(b), slope=_np(sl)))
        return plain, bias, prelu, bias_prelu

    for name, dt in (('int8', np.int8), ('int16', np.int16), ('int32', np.int32)):
        p, b, pr, bp = _rq(dt)
        setattr(ext, 'requant_to_' + name, p)
        setattr(ext, 'bias_requant_to_' + name, b)
        setattr(ext, 'prelu_requant_to_' + name, pr)
        setattr(ext, 'bias_prelu_requant_to_' + name, bp)

    def prelu(inp, slope):
        return torch.from_numpy(K.prelu(_np(inp), _np(slope)))

    class GPUHashTable:
        """hashmap_cuda.cuh:87-145 seen through its two users (cuda_ops.py:117-129): the table remembers the
        inserted coordinates (stored in the caller's key tensor is not needed on the CPU)."""
        _store = {}

        def __init__(self, keys, vals):
            self.id = keys.data_ptr()

        def insert_coords(self, coords):  # (x, y, z, batch)
            GPUHashTable._store[self.id] = coords[:, [3, 0, 1, 2]].numpy().copy()

        def lookup_coords(self, coords, kernel_sizes, strides, kernel_volume):
            inc = GPUHashTable._store[self.id]
            out = coords[:, [3, 0, 1, 2]].numpy()
            table = K.lookup_coords(inc, out, kernel_sizes.tolist(), strides.tolist())  # [K, N]
            n = out.shape[0]
            n_pad = (n + 127) // 128 * 128  # hashmap_cuda.cuh:102,339-348
            full = np.zeros((n_pad, kernel_volume), dtype=np.int32)
            full[:n] = table.T
            return torch.from_numpy(full)

    ext.cutlass_gemm_int8 = cutlass_gemm_int8
    ext.cutlass_gather_gemm_scatter_int8 = cutlass_gather_gemm_scatter_int8
    ext.softmax_int32 = softmax_int32
    ext.prelu = prelu
    ext.GPUHashTable = GPUHashTable
    return ext


# ----------------------------------------------------------------------------------------------
# import the reference with the absent third-party modules stubbed (shared with tests/test_gpu_dropin.py)
# ----------------------------------------------------------------------------------------------
from tests import ref_import  # noqa: E402


def import_reference_model():
    simple_rans = build_ref.load_ref('simple_rans_ext_cpp')
    assert simple_rans is not None, 'run python oracle/build_ref.py first'
    return ref_import.import_reference_model(REF, _make_ext(), simple_rans=simple_rans, rans=build_ref.load_ref('rans_ext_cpp'),
                                             stub_cuda_sync=True)


def build_reference_model(ref, cfg):
    sd_np = synth.make_lossl_int_state_dict(seed=7, **{k: v for k, v in cfg.items() if k != 'skip_top_scales_num'})
    return ref_import.build_reference_model(ref, cfg, sd_np)


def main():
    ref = import_reference_model()
    out = {'_doc': 'bitstreams of the reference Python codec (see make_int_codec_golden.py); bytes are hex',
           'cases': []}
    for case in CASES:
        m = build_reference_model(ref, case['cfg'])
        xyz = case_cloud(case)
        with torch.no_grad():
            data = m.compress(torch.from_numpy(synth.with_batch(xyz)))
            rec = m.decompress(data).numpy()
        assert (np.unique(rec, axis=0) == np.unique(xyz, axis=0)).all(), 'reference round trip is not lossless?'
        out['cases'].append({
            'name': case['name'], 'n_points': int(xyz.shape[0]), 'n_bytes': len(data),
            'bitstream_hex': data.hex() if len(data) <= 4096 else None,
            'bitstream_sha256': hashlib.sha256(data).hexdigest(),
            'decoded_sha256': hashlib.sha256(np.ascontiguousarray(rec.astype('<i4')).tobytes()).hexdigest(),
        })
        print(case['name'], xyz.shape[0], 'pts ->', len(data), 'bytes')
    with open(osp.join(HERE, 'int_codec_golden.json'), 'w') as f:
        json.dump(out, f, indent=1)

    # PTQ import formulas (cuda_ops.py:223-301, 464-468, 488-501, 542-600), reference classes on seeded floats
    from tests.golden.ptq_cases import run_cases
    ref_ops = importlib.import_module('lib.int_sparse_conv.cuda_ops')
    with open(osp.join(HERE, 'ptq_golden.json'), 'w') as f:
        json.dump(run_cases(ref_ops), f, indent=1)
    print('ptq cases written')

    # the convert pipeline (lossl_coord/model.py:685-888): reference functions on the reference's float trees, with
    # a parameter-only stand-in for torchsparse.nn.Conv3d (kernel [K,Cin,Cout], bias, sizes -- all the pass reads)
    import math
    import torchsparse.nn as tsn

    class Conv3dParams(nn.Module):
        def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilation=1, bias=False):
            super().__init__()
            t = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v,) * 3  # noqa: E731
            self.in_channels, self.out_channels = in_channels, out_channels
            self.kernel_size, self.stride = t(kernel_size), t(stride)
            self.kernel = nn.Parameter(torch.zeros(math.prod(self.kernel_size), in_channels, out_channels))
            self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
    tsn.Conv3d = Conv3dParams
    sys.modules['torchsparse.nn.functional'] = types.ModuleType('torchsparse.nn.functional')
    tsn.functional = sys.modules['torchsparse.nn.functional']
    fm = importlib.import_module('models.convolutional.lossl_coord.model')
    fm.spnn.Conv3d = Conv3dParams
    from tests.golden.ptq_cases import run_pipeline
    ns = types.SimpleNamespace(
        OneScalePredictor=fm.OneScalePredictor, OneScaleMultiStepPredictor=fm.OneScaleMultiStepPredictor,
        insert_obs_into_resblocks=fm.insert_obs_into_resblocks, insert_obs_into_seqs=fm.insert_obs_into_seqs,
        replace_resblocks_with_int_impl=fm.replace_resblocks_with_int_impl, replace_seqs_with_int_impl=fm.replace_seqs_with_int_impl,
        SparseTensorHistogramObserver=ref_ops.SparseTensorHistogramObserver, SparseTensor=sys.modules['torchsparse'].SparseTensor)
    with open(osp.join(HERE, 'ptq_pipeline_golden.json'), 'w') as f:
        json.dump(run_pipeline(ns), f, indent=1)
    print('ptq pipeline written')


if __name__ == '__main__':
    main()
