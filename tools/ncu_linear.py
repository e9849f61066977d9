"""ncu target: a few launches of the fused linear kernel in the regimes of a converted model.
usage (under ncu): python tools/ncu_linear.py [rows=391612]"""
import sys
import os.path as osp
import numpy as np
import torch
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from fastpcc_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 391612
ch = 256
rng = np.random.default_rng(0)
f = torch.from_numpy(rng.integers(-128, 128, (n, ch)).astype(np.int8)).cuda()
w2 = torch.from_numpy(rng.integers(-127, 128, (ch, ch)).astype(np.int8)).cuda()
bias = torch.from_numpy(rng.integers(-5000, 5000, ch).astype(np.int32)).cuda()
zp = torch.zeros(1, dtype=torch.int64, device='cuda')
slope = torch.tensor([1 << 23], dtype=torch.int32, device='cuda')
mul_hi = torch.from_numpy(rng.integers(1 << 29, 1 << 30, ch).astype(np.int64)).to(torch.uint32).cuda()
mul_lo = torch.from_numpy(rng.integers(1 << 20, 1 << 21, ch).astype(np.int64)).to(torch.uint32).cuda()
cases = [ops.make_epilogue(mul_hi, zp, 38, ops.OUT_I8, bias=bias, slope=slope),
         ops.make_epilogue(mul_hi, zp, 38, ops.OUT_I8, bias=bias),
         ops.make_epilogue(mul_lo, zp, 12, ops.OUT_I32, bias=bias)]
for _ in range(2):
    for e in cases:
        ops.linear(f, w2, e)
torch.cuda.synchronize()
