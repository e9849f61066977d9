"""Training slice of the sparse-convolution path (reference: train.py:139,215,359-404 -- a model wrapped
in DistributedDataParallel, loss.backward(), optimiser step, one process per GPU, NCCL carrying only the gradient
all-reduce).

`OccupancyNet` is the float occupancy predictor of one scale of the lossless coordinate codec
(models/convolutional/lossl_coord/model.py:137-175 in miniature): an embedding conv, residual `Block`s
(model.py:645-660) and a linear head whose 255 logits are trained with the cross entropy of the node's occupancy symbol
-- the codec's bits-per-node objective.  Every convolution is differentiable through `fastpcc_b200/autograd.py`:
forward and dgrad on the fused tcgen05 kernel, wgrad on `fpcc_spconv_wgrad_f16`.
"""
from typing import List

import numpy as np
import torch
import torch.nn as nn

from . import ops, synth
from .sparse_tensor import SparseTensor
from . import torchsparse_nn as TS


class OccupancyNet(nn.Module):
    def __init__(self, channels: int = 128, n_blocks: int = 2):
        super().__init__()
        self.embed = TS.Conv3d(8, channels, 3, 1, 1, bias=True)
        self.blocks = nn.ModuleList([TS.Block(channels) for _ in range(n_blocks)])
        self.head = nn.Linear(channels, 255)

    def forward(self, x: SparseTensor) -> torch.Tensor:
        y = self.embed(x)
        y.F = torch.nn.functional.leaky_relu(y.F, 0.1)
        for b in self.blocks:
            y = b(y)
        return self.head(y.F.float())


def make_batch(seeds: List[int], level_shift: int, device) -> tuple:
    """Nodes of stride 2^level_shift of synthetic LiDAR scans (one batch index per scan), their occupancy symbol as the
    target and the parent-level occupancy bits as the input feature (what the decoder knows when it predicts the level)."""
    cs = []
    for b, s in enumerate(seeds):
        xyz = np.unique(synth.lidar_frame(s) >> (level_shift - 1), axis=0)  # children of the nodes we predict
        cs.append(synth.with_batch(xyz, b))
    child = torch.from_numpy(np.concatenate(cs)).to(device)
    order = torch.argsort(ops.morton_encode(child, col0=1, msb_axis=0) + (child[:, 0].long() << 52))
    child = ops.gather_rows(child.contiguous(), order)
    node_c, occ, _, _, cnt = ops.downsample(child)
    n = int(cnt.item())
    node_c, occ = node_c[:n].contiguous(), occ[:n].contiguous()
    # input feature: the occupancy bits of the node's PARENT (known context), broadcast to the node
    par_c, par_occ, par_of, _, cnt2 = ops.downsample(node_c)
    bits = ops.occ_to_bits(par_occ[: int(cnt2.item())].contiguous()).float()
    feats = bits[par_of[:n].long()]
    target = occ.long() - 1
    return SparseTensor(feats, node_c, 1), target


def train_step(model: nn.Module, opt: torch.optim.Optimizer, batch) -> torch.Tensor:
    x, target = batch
    opt.zero_grad(set_to_none=True)
    logits = model(SparseTensor(x.F, x.C, x.stride))
    loss = torch.nn.functional.cross_entropy(logits, target) / np.log(2.0)  # bits per node
    loss.backward()
    opt.step()
    return loss.detach()
