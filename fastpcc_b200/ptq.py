"""Post-training quantisation of the float lossless codec into the integer one: observers, the float
module trees of `lossl_coord`, and the convert pass that replaces them with the integer layers of
`fastpcc_b200/int_sparse_conv/cuda_ops.py` through their `import_parameters`.

Mirrors, with the reference's names: `make_obs`, `SparseTensorHistogramObserver`, `SparseResBlockWithObs`
(lib/int_sparse_conv/cuda_ops.py:20-59), `insert_obs_into_resblocks`, `insert_obs_into_seqs`,
`replace_resblocks_with_int_impl`, `replace_seqs_with_int_impl` (models/convolutional/lossl_coord/model.py:685-888,
driven by `pre_test_hook` / `post_test_hook`, :633-642) and the float `OneScalePredictor` /
`OneScaleMultiStepPredictor` trees (:28-47, 120-175).  The resulting integer state dict is the one
`lossl_coord_int.Model` loads.  Pinned against the reference's own functions by tests/test_ptq_golden.py."""
from typing import List, Optional

import torch
import torch.nn as nn
from torch.ao.quantization import HistogramObserver

from .int_sparse_conv.cuda_ops import (ActRange, LinearIn8W8Out8, LinearIn8W8Out32, LinearPReLUIn8W8Out8,
                                       LinearPReLUIn8W8Out32, PReLUIn32Out32, RequantFxpToScaledInt8,
                                       SparseConvIn8W8Out8, SparseConvIn8W8Out32, SparseConvPReLUIn8W8Out8,
                                       SparseConvPReLUIn8W8Out32, SparseResBlockIn32W8Out32)
from .sparse_tensor import SparseTensor
from .torchsparse_nn import Block, Conv3d, SparseSequential


def _int_sequential(*mods):
    """the integer model's SparseSequential (dense layers take features, sparse layers take SparseTensors)"""
    from .lossl_coord_int.model import SparseSequential as IntSequential
    return IntSequential(*mods)


class SparseTensorHistogramObserver(HistogramObserver):
    """cuda_ops.py:20-32: observes the features of the SparseTensor passing through."""

    def forward(self, input: SparseTensor) -> SparseTensor:
        super().forward(input.F.detach().float())
        return input

    def extra_repr(self):
        return f'min_val={self.min_val}, max_val={self.max_val}, {self.qscheme}'


def make_obs(qscheme=torch.per_tensor_symmetric):
    return SparseTensorHistogramObserver(bins=2048, dtype=torch.qint8, quant_min=-ActRange, quant_max=ActRange, qscheme=qscheme)


class SparseResBlockWithObs(nn.Module):
    """cuda_ops.py:40-59: the float residual block with an observer in front of each conv."""

    def __init__(self, ch):
        super().__init__()
        self.ch = ch
        self.obs = make_obs(torch.per_tensor_symmetric)
        self.conv = Conv3d(ch, ch, 3, 1, 1, bias=True)
        self.act = nn.PReLU()
        self.obs2 = make_obs(torch.per_tensor_symmetric)
        self.conv2 = Conv3d(ch, ch, 3, 1, 1, bias=True)
        self.act2 = nn.PReLU()

    def forward(self, org: SparseTensor) -> SparseTensor:
        org = self.obs(org)
        x = self.conv(org)
        x.F = self.act(x.F.to(self.act.weight.dtype))
        x = self.obs2(x)
        x = self.conv2(x)
        x.F = self.act2(x.F.to(self.act2.weight.dtype) + org.F.to(self.act2.weight.dtype))
        return x


# ---- float module trees (lossl_coord/model.py:28-47, 120-175): what a float checkpoint of the reference loads into ----

class OneScalePredictor(nn.Module):
    def __init__(self, channels, if_upsample=True, allow_single_ch=False):
        super().__init__()
        if allow_single_ch:
            self.dec_init = Conv3d(1, channels, 3, 1, 1, bias=True)
        self.dec = Block(channels)
        self.pred = SparseSequential(Conv3d(channels, channels, 3, 1, 1, bias=True), nn.PReLU(), nn.Linear(channels, 255))
        self.if_upsample = if_upsample
        self.upsample = SparseSequential(nn.Linear(channels + 8, channels), nn.PReLU(), Block(channels),
                                         nn.Linear(channels, channels * 8)) if if_upsample else None


class OneScaleMultiStepPredictor(nn.Module):
    def __init__(self, channels, pred_steps=2, use_more_ch_for_multi_step_pred=True):
        super().__init__()
        self.pred_steps = pred_steps
        k = 2 ** (pred_steps - 2)
        if pred_steps == 2:
            self.embed, out_ch, cin = SparseSequential(), channels, channels + 8
        elif use_more_ch_for_multi_step_pred:
            emb = 64 if pred_steps == 3 else 512
            self.embed = SparseSequential(Conv3d(8, emb, 2 if pred_steps == 3 else k, 2 if pred_steps == 3 else k, bias=True), nn.PReLU())
            out_ch = round(channels * 1.25) if pred_steps == 3 else channels * 2
            cin = (channels if pred_steps == 3 else round(channels * 1.25)) + emb
        else:
            assert pred_steps >= 3
            self.embed = SparseSequential(Conv3d(8, channels, k, k, bias=True))
            if channels >= 256:
                self.embed.append(nn.PReLU())
            out_ch, cin = channels, channels * 2
        self.dec = SparseSequential(nn.Linear(cin, out_ch), nn.PReLU(), Block(out_ch)) if cin != out_ch else Block(out_ch)
        self.pred = nn.ModuleList()
        for idx in range(pred_steps):
            if idx == 0:
                self.pred.append(SparseSequential(Conv3d(out_ch, out_ch, 3, 1, 1, bias=True), nn.PReLU(), nn.Linear(out_ch, channels * 8)))
            elif idx != pred_steps - 1:
                self.pred.append(SparseSequential(nn.PReLU(), nn.Linear(channels + 8, channels), nn.PReLU(),
                                                  Conv3d(channels, channels, 3, 1, 1, bias=True), nn.PReLU(),
                                                  nn.Linear(channels, channels * 8)))
            else:
                self.pred.append(SparseSequential(Conv3d(channels, channels, 3, 1, 1, bias=True), nn.PReLU(), nn.Linear(channels, 255)))


class FloatTrunk(nn.Module):
    """The parameter tree of the float lossl_coord Model (model.py:228-238): `blocks_dec.*`, `block_dec_recurrent.*`."""

    def __init__(self, channels=256, max_stride_wo_recurrent=2048, max_stride=8192, fea_stride=16,
                 use_more_ch_for_multi_step_pred=False):
        super().__init__()
        import math
        self.blocks_dec = nn.ModuleList()
        for idx in range(int(math.log2(max_stride_wo_recurrent))):
            steps = int(math.log2(fea_stride)) - idx
            if steps < 1:
                self.blocks_dec.append(OneScalePredictor(channels, True, False))
            elif steps == 1:
                self.blocks_dec.append(OneScalePredictor(channels, False, False))
            else:
                self.blocks_dec.append(OneScaleMultiStepPredictor(channels, steps, use_more_ch_for_multi_step_pred))
        self.block_dec_recurrent = OneScalePredictor(channels, True, True)


# ---- observer insertion (model.py:685-722) -------------------------------------------------------------------------

def _device_of(model):
    p = next(model.parameters(), None)
    return p.device if p is not None else torch.device('cpu')


def _walk(parent: nn.Module, visit):
    """`visit(parent, name, child)` returns True when it replaced / consumed the child (no descent into it)."""
    for name, child in list(parent._modules.items()):
        if child is None or visit(parent, name, child):
            continue
        _walk(child, visit)


def insert_obs_into_resblocks(model: nn.Module):
    dev = _device_of(model)

    def visit(parent, name, child):
        if not isinstance(child, Block):
            return False
        new = SparseResBlockWithObs(child.ch).to(dev)
        new.load_state_dict(child.state_dict(), strict=False)
        parent._modules[name] = new
        return True
    _walk(model, visit)


def insert_obs_into_seqs(model: nn.Module):
    """An observer in front of every element of a non-empty SparseSequential: affine (zero point allowed) when the
    element it feeds is an nn.Linear, symmetric otherwise."""
    dev = _device_of(model)

    def scheme(m):
        return torch.per_tensor_affine if isinstance(m, nn.Linear) else torch.per_tensor_symmetric

    def visit(parent, name, child):
        if not (isinstance(child, SparseSequential) and len(child) > 0):
            return False
        mods = list(child)
        new: List[nn.Module] = []
        for m in mods:
            new += [make_obs(scheme(m)).to(dev), m]
        parent._modules[name] = SparseSequential(*new)
        return True
    _walk(model, visit)


# ---- conversion (model.py:725-888) ---------------------------------------------------------------------------------

def replace_resblocks_with_int_impl(model: nn.Module):
    dev = _device_of(model)

    def visit(parent, name, child):
        if not isinstance(child, SparseResBlockWithObs):
            return False
        new = SparseResBlockIn32W8Out32(child.ch).to(dev)
        new.import_parameters(child)
        parent._modules[name] = new
        return True
    _walk(model, visit)


_AFFINE = (nn.Linear, Conv3d)
# (is_linear, fused PReLU, output feeds another affine layer as scaled int8) -> integer layer class
_INT_CLASS = {
    (True, True, True): LinearPReLUIn8W8Out8, (True, True, False): LinearPReLUIn8W8Out32,
    (True, False, True): LinearIn8W8Out8, (True, False, False): LinearIn8W8Out32,
    (False, True, True): SparseConvPReLUIn8W8Out8, (False, True, False): SparseConvPReLUIn8W8Out32,
    # the reference instantiates SparseConvIn8W8Out32 here and then passes it the Out8 argument list (a TypeError,
    # model.py:849-851); no codec topology reaches that branch.  The consistent choice is the Out8 layer.
    (False, False, True): SparseConvIn8W8Out8, (False, False, False): SparseConvIn8W8Out32,
}


def _convert_sequential(seq: SparseSequential, dev) -> SparseSequential:
    items = list(seq)
    is_obs = [isinstance(m, SparseTensorHistogramObserver) for m in items]
    layers = [i for i, o in enumerate(is_obs) if not o]  # positions of the real modules

    def obs_before(pos) -> SparseTensorHistogramObserver:
        return next(items[k] for k in range(pos - 1, -1, -1) if is_obs[k])

    def obs_after(pos) -> SparseTensorHistogramObserver:
        return next(items[k] for k in range(pos + 1, len(items)) if is_obs[k])

    out: List[nn.Module] = []
    scaled_int = False  # is the running activation a scaled int8 (True) or Q8.23 fixed point (False)
    j = 0
    while j < len(layers):
        pos = layers[j]
        m = items[pos]
        if isinstance(m, _AFFINE):
            prelu = items[layers[j + 1]] if j + 1 < len(layers) and isinstance(items[layers[j + 1]], nn.PReLU) else None
            nxt = j + (2 if prelu is not None else 1)
            out8 = nxt < len(layers) and isinstance(items[layers[nxt]], _AFFINE)
            scale_in, zp_in = obs_before(pos).calculate_qparams()
            if not scaled_int:
                rq = RequantFxpToScaledInt8().to(dev)
                rq.import_parameters(scale_in, zp_in)
                out.append(rq)
            is_lin = isinstance(m, nn.Linear)
            cls = _INT_CLASS[(is_lin, prelu is not None, out8)]
            layer = (cls(m.in_features, m.out_features) if is_lin else cls(m.in_channels, m.out_channels, m.kernel_size, m.stride)).to(dev)
            args = [scale_in, zp_in]
            if out8:
                args += list(obs_after(layers[nxt - 1]).calculate_qparams())
            args.append(m)
            if prelu is not None:
                args.append(prelu)
            layer.import_parameters(*args)
            out.append(layer)
            scaled_int = out8
            j = nxt
        elif isinstance(m, nn.PReLU):
            if scaled_int:
                raise NotImplementedError('a stand-alone PReLU on a scaled-int8 activation')
            p = PReLUIn32Out32().to(dev)
            p.import_parameters(m)
            out.append(p)
            j += 1
        elif isinstance(m, SparseResBlockIn32W8Out32):
            if scaled_int:
                raise NotImplementedError('a residual block on a scaled-int8 activation')
            out.append(m)
            j += 1
        else:
            raise NotImplementedError(m)
    return _int_sequential(*out)


def replace_seqs_with_int_impl(model: nn.Module):
    dev = _device_of(model)

    def visit(parent, name, child):
        if isinstance(child, SparseSequential) and not type(child).__module__.endswith('lossl_coord_int.model'):
            parent._modules[name] = _convert_sequential(child, dev)
            return True
        if isinstance(child, Conv3d) and name.endswith('dec_init'):  # the one conv outside a sequential: input is the ones feature
            new = SparseConvIn8W8Out32(1, child.out_channels, child.kernel_size, child.stride).to(dev)
            new.import_parameters(torch.tensor((1,), dtype=torch.float32, device=dev), torch.tensor((0,), dtype=torch.int32, device=dev), child)
            parent._modules[name] = new
            return True
        return False
    _walk(model, visit)


def insert_observers(model: nn.Module):
    """`pre_test_hook` with quantize_param (model.py:633-636)"""
    insert_obs_into_resblocks(model)
    insert_obs_into_seqs(model)


def convert_to_int(model: nn.Module, save_path: Optional[str] = None):
    """`post_test_hook` (model.py:638-642): after calibration data has passed through the observers."""
    replace_resblocks_with_int_impl(model)
    replace_seqs_with_int_impl(model)
    if save_path:
        torch.save({'state_dict': model.state_dict()}, save_path)
    return model
