"""Seeded float layers for the PTQ-import golden vectors (tests/golden/ptq_golden.json): each case calls
`import_parameters` of one integer layer class with the arguments the reference's convert step passes
(lossl_coord/model.py:685-888 -> cuda_ops.py:223-301, 464-468, 488-501, 542-600)."""
import hashlib
import types

import numpy as np
import torch
import torch.nn as nn


def _conv(g, cin, cout, ks, stride):
    kv = ks[0] * ks[1] * ks[2]
    c = types.SimpleNamespace()
    c.kernel = torch.randn(kv, cin, cout, generator=g) * 0.05
    c.bias = torch.randn(cout, generator=g) * 0.1
    c.kernel_size, c.stride = ks, stride
    return c


def _linear(g, cin, cout):
    lin = nn.Linear(cin, cout)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(cout, cin, generator=g) * 0.08)
        lin.bias.copy_(torch.randn(cout, generator=g) * 0.1)
    return lin


def _prelu(v):
    p = nn.PReLU()
    with torch.no_grad():
        p.weight.fill_(v)
    return p


def _t(v, dt=torch.float32):
    return torch.tensor([v], dtype=dt)


def run_cases(mod):
    """`mod`: a module exposing the layer classes of lib/int_sparse_conv/cuda_ops.py.  Returns
    {case name: {buffer name: sha256 of the little-endian bytes}} for the integer (persistent) buffers."""
    g = torch.Generator().manual_seed(1234)
    out = {}

    def record(name, layer):
        d = {}
        for k, v in layer.state_dict().items():
            if k.split('.')[-1].startswith(('scale_', 'zero_point_')):
                continue  # float bookkeeping of the calibration, not read by the integer path
            a = v.view(torch.int32).numpy() if v.dtype == torch.uint32 else v.numpy()
            d[k] = hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() + ':' + str(a.dtype) + str(list(a.shape))
        out[name] = d

    zi = lambda v: _t(v, torch.int64)  # noqa: E731
    layer = mod.SparseConvIn8W8Out8(16, 24)
    layer.import_parameters(_t(0.031), zi(0), _t(0.017), zi(0), _conv(g, 16, 24, (3, 3, 3), (1, 1, 1)))
    record('conv_out8', layer)
    layer = mod.SparseConvIn8W8Out32(8, 32, (2, 2, 2), (2, 2, 2))
    layer.import_parameters(_t(1.0), zi(0), _conv(g, 8, 32, (2, 2, 2), (2, 2, 2)))
    record('conv_out32_stride2', layer)
    layer = mod.SparseConvPReLUIn8W8Out8(32, 32)
    layer.import_parameters(_t(0.0123), zi(3), _t(0.02), zi(-5), _conv(g, 32, 32, (3, 3, 3), (1, 1, 1)), _prelu(0.25))
    record('conv_prelu_out8_zero_points', layer)
    layer = mod.SparseConvPReLUIn8W8Out32(16, 16)
    layer.import_parameters(_t(0.05), zi(0), _conv(g, 16, 16, (3, 3, 3), (1, 1, 1)), _prelu(0.1))
    record('conv_prelu_out32', layer)
    layer = mod.LinearIn8W8Out8(40, 16)
    layer.import_parameters(_t(0.02), zi(0), _t(0.04), zi(0), _linear(g, 40, 16))
    record('linear_out8', layer)
    layer = mod.LinearIn8W8Out32(32, 255)
    layer.import_parameters(_t(0.015), zi(0), _linear(g, 32, 255))
    record('linear_out32_logits', layer)
    layer = mod.LinearPReLUIn8W8Out8(24, 24)
    layer.import_parameters(_t(0.03), zi(2), _t(0.025), zi(1), _linear(g, 24, 24), _prelu(0.2))
    record('linear_prelu_out8_zero_points', layer)
    layer = mod.LinearPReLUIn8W8Out32(16, 128)
    layer.import_parameters(_t(0.02), zi(0), _linear(g, 16, 128), _prelu(0.3))
    record('linear_prelu_out32', layer)
    layer = mod.RequantFxpToScaledInt8()
    layer.import_parameters(_t(0.0217), zi(0))
    record('requant_fxp', layer)
    layer = mod.RequantFxpToScaledInt8()
    layer.import_parameters(_t(0.4), zi(-7))
    record('requant_fxp_zero_point', layer)
    layer = mod.PReLUIn32Out32()
    layer.import_parameters(_prelu(0.2))
    record('prelu32', layer)
    return out


# ---- the convert pipeline (lossl_coord/model.py:685-888) on the float predictor trees ---------------------------

def seed_parameters(model):
    """Deterministic values by parameter name (identical for the reference tree and the twin)."""
    import zlib
    for name, p in model.named_parameters():
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
        with torch.no_grad():
            if p.numel() == 1:  # PReLU slope
                p.fill_(0.05 + 0.25 * torch.rand(1, generator=g).item())
            else:
                p.copy_(torch.randn(p.shape, generator=g) * (0.1 if p.dim() == 1 else 0.05))


def run_pipeline(ns):
    """`ns`: namespace with OneScalePredictor, OneScaleMultiStepPredictor, the two insert_* and two replace_*
    functions, SparseTensorHistogramObserver and SparseTensor.  Builds float trees, inserts observers, feeds every
    observer seeded activations, converts, and returns the hashes of the integer state dict."""
    import zlib
    model = nn.ModuleDict({
        'recurrent': ns.OneScalePredictor(16, True, True),
        'plain': ns.OneScalePredictor(24, False, False),
        'two_step': ns.OneScaleMultiStepPredictor(16, 2, False),
        'three_step': ns.OneScaleMultiStepPredictor(16, 3, False),
        'four_step_wide': ns.OneScaleMultiStepPredictor(16, 4, True),
        'three_step_256': ns.OneScaleMultiStepPredictor(256, 3, False),
    })
    seed_parameters(model)
    ns.insert_obs_into_resblocks(model)
    ns.insert_obs_into_seqs(model)
    for name, m in model.named_modules():
        if isinstance(m, ns.SparseTensorHistogramObserver):
            g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
            spread = 0.2 + 3.0 * torch.rand(1, generator=g).item()
            shift = 0.4 if m.qscheme == torch.per_tensor_affine else 0.0
            for _ in range(2):
                m(ns.SparseTensor(torch.randn(600, 8, generator=g) * spread + shift, torch.zeros((600, 4), dtype=torch.int32)))
    ns.replace_resblocks_with_int_impl(model)
    ns.replace_seqs_with_int_impl(model)
    out = {}
    for k, v in model.state_dict().items():
        if k.split('.')[-1].startswith(('scale_', 'zero_point_')):
            continue
        a = v.view(torch.int32).numpy() if v.dtype == torch.uint32 else v.numpy()
        out[k] = hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:24] + ':' + str(a.dtype) + str(list(a.shape))
    return out
