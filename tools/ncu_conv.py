"""ncu target: the two sparse-conv flavours of a ResBlock on one LiDAR level (grouped rows), a few launches each.
usage (under ncu): python tools/ncu_conv.py [stride_log2=4] [frames=8]"""
import sys
import os.path as osp
import numpy as np
import torch
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from fastpcc_b200 import ops, synth  # noqa: E402

lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 4
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ch = 256
cs = []
for b in range(frames):
    xyz = synth.lidar_frame(1000 + b) >> lvl
    cs.append(synth.with_batch(np.unique(xyz, axis=0), b))
C = torch.from_numpy(np.concatenate(cs)).cuda()
n = C.shape[0]
rng = np.random.default_rng(0)
f = torch.from_numpy(rng.integers(-128, 128, (n, ch)).astype(np.int8)).cuda()
w = torch.from_numpy(rng.integers(-127, 128, (27, ch, ch)).astype(np.int8)).cuda()
bias = torch.from_numpy(rng.integers(-5000, 5000, ch).astype(np.int32)).cuda()
zp = torch.zeros(1, dtype=torch.int64, device='cuda')
slope = torch.tensor([1 << 23], dtype=torch.int32, device='cuda')
mul_hi = torch.from_numpy(rng.integers(1 << 29, 1 << 30, ch).astype(np.int64)).to(torch.uint32).cuda()
mul_lo = torch.from_numpy(rng.integers(1 << 20, 1 << 21, ch).astype(np.int64)).to(torch.uint32).cuda()
res = torch.from_numpy(rng.integers(-(1 << 28), 1 << 28, (n, ch)).astype(np.int32)).cuda()
keys, vals = ops.hash_build(C)
table = ops.kmap_lookup(keys, vals, C, (3, 3, 3), (1, 1, 1))
tp, perm = ops.group_rows(table)
ep8 = ops.make_epilogue(mul_hi, zp, 38, ops.OUT_I8, bias=bias, slope=slope)
ep32 = ops.make_epilogue(mul_lo, zp, 12, ops.OUT_I32, bias=bias, residual=res, post_slope=slope)
for _ in range(2):
    ops.spconv(f, w, tp, ep8, row_perm=perm)
    ops.spconv(f, w, tp, ep32, row_perm=perm)
torch.cuda.synchronize()
