"""Imports the reference's OWN `models/convolutional/lossl_coord_int/model.py` (+ `lib/int_sparse_conv/cuda_ops.py`),
unmodified, with the absent third-party modules stood in.  TEST INFRASTRUCTURE ONLY; used by

  * tests/golden/make_int_codec_golden.py  (this container, CPU: `ext` = numpy stand-in of the CUDA extension)
  * tests/test_gpu_dropin.py               (B200: `ext` = fastpcc_b200.int_sparse_conv.ext, the product's drop-in)

`ref_root` is /root/reference here, or a temporary extraction of oracle/_ref/pyref.zip on the GPU box (a git-ignored
archive of the reference's .py files staged by oracle/build_ref.py, like a `pip install --target` of the reference:
it travels with the snapshot, it is never committed).
"""
import importlib
import os.path as osp
import sys
import types

import torch
import torch.nn as nn


def import_reference_model(ref_root, ext, sparse_tensor_cls=None, morton_ext=None, simple_rans=None, rans=None,
                           stub_cuda_sync=False):
    """-> the imported module `models.convolutional.lossl_coord_int.model` of the reference tree at `ref_root`.

    ext                pybind-module stand-in bound as lib.int_sparse_conv.build.int_sparse_conv_ext (binding.cu:114-145)
    sparse_tensor_cls  class bound as torchsparse.SparseTensor (torchsparse is not installable offline)
    morton_ext         module bound as the result of the space_filling_curves JIT load (morton3d.cu), or None
    simple_rans, rans  the reference's compiled range coders (oracle/_ref), returned by the stubbed JIT loader
    """
    if ref_root not in sys.path:
        sys.path.insert(1, ref_root)
    ts = types.ModuleType('torchsparse')
    tsn = types.ModuleType('torchsparse.nn')

    if sparse_tensor_cls is None:
        class SparseTensor:
            def __init__(self, feats, coords, stride=(1, 1, 1), spatial_range=None):
                self.F, self.C, self.spatial_range = feats, coords, spatial_range
                self.stride = tuple(stride) if isinstance(stride, (tuple, list)) else (stride,) * 3
                self._caches = types.SimpleNamespace(cmaps={}, kmaps={}, hashmaps={})
        sparse_tensor_cls = SparseTensor

    class Conv3d(nn.Module):
        pass

    ts.SparseTensor, ts.nn, tsn.Conv3d = sparse_tensor_cls, tsn, Conv3d
    sys.modules['torchsparse'], sys.modules['torchsparse.nn'] = ts, tsn
    for name in ('plyfile', 'open3d', 'cv2'):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                m = types.ModuleType(name)
                m.PlyData = m.PlyElement = object
                sys.modules[name] = m

    # compiled-extension loaders: hand back what is prebuilt from the reference's own C++ (range coders) or the given
    # stand-ins; nothing is compiled into the reference tree.
    import torch.utils.cpp_extension as cpp_ext

    def fake_load(name, *a, **kw):
        if name == 'simple_rans_ext_cpp':
            return simple_rans
        if name == 'rans_ext_cpp':
            return rans
        if name == 'space_filling_curves_ext' and morton_ext is not None:
            return morton_ext
        return types.ModuleType(name)  # CUDA-only helpers (knn ...) never reached on this path
    cpp_ext.load = fake_load

    build_mod = types.ModuleType('lib.int_sparse_conv.build')
    build_mod.int_sparse_conv_ext = ext
    sys.modules['lib.int_sparse_conv.build'] = build_mod

    # lossy_coord_v3/__init__ pulls in the whole v3 model (torchsparse internals); only its rans_coder sub-package is
    # needed: register the package without executing its __init__.
    pkg = types.ModuleType('models.convolutional.lossy_coord_v3')
    pkg.__path__ = [osp.join(ref_root, 'models/convolutional/lossy_coord_v3')]
    sys.modules['models.convolutional.lossy_coord_v3'] = pkg

    if stub_cuda_sync:
        torch.cuda.synchronize = lambda *a, **k: None
    for _ in range(20):
        try:
            return importlib.import_module('models.convolutional.lossl_coord_int.model')
        except ModuleNotFoundError as e:  # optional third-party imports of unrelated helpers
            if e.name.split('.')[0] in ('lib', 'models'):
                raise
            sys.modules[e.name] = types.ModuleType(e.name)
    raise RuntimeError('could not import the reference model')


def build_reference_model(ref, cfg, sd_np, device='cpu'):
    """reference Model(cfg) loaded with a {key: ndarray} state dict (uint32 multipliers travel as int32 views)"""
    import numpy as np
    c = ref.Config()
    for k, v in cfg.items():
        setattr(c, k, v)
    m = ref.Model(c, torch.device(device))
    sd = {}
    for k, v in sd_np.items():
        t = torch.from_numpy(np.ascontiguousarray(v.view(np.int32) if v.dtype == np.uint32 else v))
        sd[k] = t.view(torch.uint32) if v.dtype == np.uint32 else t
    own = m.state_dict()
    missing = [k for k in own if k not in sd and not k.split('.')[-1].startswith(('scale_', 'zero_point_'))]
    assert not missing, missing
    m.load_state_dict(sd, strict=False)
    m.eval()
    return m.to(device)


LOSSY_V2_BASELINE_R1 = dict(  # config/convolutional/lossy_coord_v2/baseline_r1.yaml:2-14
    activation='prelu', compressed_channels=(1,), skip_encoding_fea=1, encoder_channels=(16, 64), decoder_channels=(16,),
    adaptive_pruning=True, geo_lossl_if_sample=(0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1),
    geo_lossl_channels=(64, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 1),
    bits_loss_factor=0.4, warmup_fea_loss_steps=5000, warmup_fea_loss_factor=0.01)


def import_reference_lossy_v2(ref_root, me_module, rans=None, morton_ext=None):
    """-> the reference's UNMODIFIED `models.convolutional.lossy_coord_v2.model` (PCC: Encoder, GeoLosslessEntropyModel,
    Decoder with top-k pruning; lossy_coord_v2/model.py:230-275, layers.py, lossy_coord_lossy_color/geo_lossl_em.py) and
    its own lib/minkowski_sparse_conv_layers.py, running on `me_module` bound as `MinkowskiEngine`
    (fastpcc_b200.me on the GPU, oracle.me_cpu.load() on the CPU).  `rans`: the reference's compiled rans_ext_cpp."""
    if ref_root not in sys.path:
        sys.path.insert(1, ref_root)
    sys.modules['MinkowskiEngine'] = me_module
    sub = types.ModuleType('MinkowskiEngine.MinkowskiSparseTensor')
    sub.SparseTensorQuantizationMode = me_module.SparseTensorQuantizationMode
    sub.SparseTensor = me_module.SparseTensor
    sys.modules['MinkowskiEngine.MinkowskiSparseTensor'] = sub
    for name in ('plyfile', 'open3d', 'cv2'):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                m = types.ModuleType(name)
                m.PlyData = m.PlyElement = object
                sys.modules[name] = m
    import torch.utils.cpp_extension as cpp_ext

    def fake_load(name, *a, **kw):
        if name == 'rans_ext_cpp':
            return rans
        if name == 'space_filling_curves_ext' and morton_ext is not None:
            return morton_ext
        return types.ModuleType(name)
    cpp_ext.load = fake_load
    # drop modules imported earlier against another backend (the importer may be used twice in one process)
    for k in [k for k in sys.modules if k.startswith(('models.convolutional.lossy_coord', 'lib.minkowski_sparse_conv_layers'))]:
        del sys.modules[k]
    for _ in range(20):
        try:
            return importlib.import_module('models.convolutional.lossy_coord_v2.model')
        except ModuleNotFoundError as e:
            if e.name.split('.')[0] in ('lib', 'models'):
                raise
            sys.modules[e.name] = types.ModuleType(e.name)
    raise RuntimeError('could not import the reference lossy_coord_v2 model')


def build_reference_lossy_v2(ref, cfg_overrides, seed=0, device='cpu'):
    """PCC(ModelConfig(**overrides)) with seeded parameters (there is no checkpoint offline): default torch init under
    `seed`, eval mode."""
    torch.manual_seed(seed)
    cfg = ref.ModelConfig()
    for k, v in cfg_overrides.items():
        setattr(cfg, k, v)
    cfg.check_local_value()
    model = ref.PCC(cfg)
    # Seeded Kaiming-uniform weights and small random biases: with the shim's default init (or zero biases) the
    # activations shrink layer by layer and every rounded residual is 0, a histogram the reference's own coder rejects
    # (cdf_ops.cpp:29).  Generated on the CPU so that both backends get identical parameters.
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if 'bottom_fea_entropy_model' in name:
                continue
            if p.dim() >= 2 and p.shape[-2] * (p.shape[0] if p.dim() == 3 else 1) > 1:
                fan_in = p.shape[-2] * (p.shape[0] if p.dim() == 3 else 1) if name.endswith('kernel') else p.shape[-1]
                bound = (6.0 / fan_in) ** 0.5
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
            elif name.endswith('bias'):
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.1)
    model.eval()
    return model.to(device)
