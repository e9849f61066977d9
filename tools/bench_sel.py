"""Micro-benchmark of the selection linear with the fused second stage (Linear(C -> 8C)[child mask] -> [PReLU] -> Requant of
the multi-step predictors): usage python tools/bench_sel.py [level=4] [frames=8]"""
import sys
import os.path as osp
import numpy as np
import torch
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from fastpcc_b200 import ops, synth  # noqa: E402

lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 4
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ch = 256
rng = np.random.default_rng(0)
cs = [synth.with_batch(np.unique(synth.lidar_frame(1000 + b) >> lvl, axis=0), b) for b in range(frames)]
fine = [synth.with_batch(np.unique(synth.lidar_frame(1000 + b) >> (lvl - 1), axis=0), b) for b in range(frames)]
Cc = torch.from_numpy(np.concatenate(cs)).cuda()
n = Cc.shape[0]
# occupancy of every coarse node from the finer level
occ = np.zeros(n, np.uint8)
key = {}
allc = np.concatenate(cs)
order = {tuple(r): i for i, r in enumerate(allc.tolist())}
for f in fine:
    par = f.copy(); par[:, 1:] >>= 1
    slot = ((f[:, 1] & 1) << 2) | ((f[:, 2] & 1) << 1) | (f[:, 3] & 1)
    idx = np.array([order[tuple(r)] for r in par.tolist()])
    np.bitwise_or.at(occ, idx, (1 << (7 - slot)).astype(np.uint8))
occ_t = torch.from_numpy(occ).cuda()
_, par, slot, n_child = ops.upsample(Cc, occ_t)
sel = ops.slot_pairs(par, slot)
a = torch.from_numpy(rng.integers(-128, 128, (n, ch)).astype(np.int8)).cuda()
w8 = torch.from_numpy(rng.integers(-127, 128, (8 * ch, ch)).astype(np.int8)).cuda()
bias8 = torch.from_numpy(rng.integers(-100000, 100000, 8 * ch).astype(np.int32)).cuda()
mul8 = torch.from_numpy(rng.integers(1 << 18, 1 << 21, 8 * ch).astype(np.int64)).to(torch.uint32).cuda()
zp0 = torch.zeros(1, dtype=torch.int64, device='cuda')
mul2 = torch.tensor([(1 << 30) + 77], dtype=torch.int64).to(torch.uint32).cuda()
sl2 = torch.tensor([int(0.3 * (1 << 25))], dtype=torch.int32, device='cuda')


def timeit(fn, it=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / it


for name, post in (('post2 prelu', (mul2, zp0, 47, sl2)), ('post2', (mul2, zp0, 47, None)), ('int32 out', None)):
    ep = ops.make_epilogue(mul8, zp0, 10, ops.OUT_I32, bias=bias8, post_requant=post)
    ms = timeit(lambda: ops.linear(a, w8, ep, sel=sel, n_out_rows=n_child))
    print(f'sel linear {name:12s} rows={n} children={n_child} 256->8x256: {ms:.3f} ms  {2.0 * n_child * ch * ch / ms / 1e9:7.1f} TOP/s')
