"""Per-tile timeline of the roles of CTA 0 of the tensor-core conv kernel (experiment build with -DFPCC_TC_TRACE).
usage: python tools/trace_tiles.py   (builds the variant here; run the printed gpurun command, or run on a GPU box)"""
import ctypes as C
import os
import os.path as osp
import sys
ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == 'build':
    from fastpcc_b200 import build
    print(build.build_variant('trace', ['-DFPCC_TC_TRACE']))
    print("gpurun -- 'FPCC_LIB_PATH=fastpcc_b200/_C/variants/trace/libfastpcc_b200.so python tools/trace_tiles.py'")
    sys.exit(0)

import numpy as np
import torch
from fastpcc_b200 import _lib, ops, synth

lib = _lib.load()
lvl, ch, frames = 4, 256, 8
cs = [synth.with_batch(np.unique(synth.lidar_frame(1000 + b) >> lvl, axis=0), b) for b in range(frames)]
Cc = torch.from_numpy(np.concatenate(cs)).cuda()
n = Cc.shape[0]
rng = np.random.default_rng(0)
f = torch.from_numpy(rng.integers(-128, 128, (n, ch)).astype(np.int8)).cuda()
w = torch.from_numpy(rng.integers(-127, 128, (27, ch, ch)).astype(np.int8)).cuda()
bias = torch.from_numpy(rng.integers(-5000, 5000, ch).astype(np.int32)).cuda()
zp = torch.zeros(1, dtype=torch.int64, device='cuda')
slope = torch.tensor([1 << 23], dtype=torch.int32, device='cuda')
mul_hi = torch.from_numpy(rng.integers(1 << 29, 1 << 30, ch).astype(np.int64)).to(torch.uint32).cuda()
mul_lo = torch.from_numpy(rng.integers(1 << 20, 1 << 21, ch).astype(np.int64)).to(torch.uint32).cuda()
res = torch.from_numpy(rng.integers(-(1 << 28), 1 << 28, (n, ch)).astype(np.int32)).cuda()
keys, vals = ops.hash_build(Cc)
table = ops.kmap_lookup(keys, vals, Cc, (3, 3, 3), (1, 1, 1))
tp, perm = ops.group_rows(table)
w2 = torch.from_numpy(rng.integers(-127, 128, (ch, ch)).astype(np.int8)).cuda()
cases = {'conv i8 prelu': lambda: ops.spconv(f, w, tp, ops.make_epilogue(mul_hi, zp, 38, ops.OUT_I8, bias=bias, slope=slope), row_perm=perm),
         'conv2 i32 res': lambda: ops.spconv(f, w, tp, ops.make_epilogue(mul_lo, zp, 12, ops.OUT_I32, bias=bias, residual=res, post_slope=slope), row_perm=perm),
         'linear i8': lambda: ops.linear(f, w2, ops.make_epilogue(mul_hi, zp, 38, ops.OUT_I8, bias=bias))}
fn = getattr(C.CDLL(_lib.so_path()), 'fpcc_trace_read')
names = ['prod start', 'prod tmem_empty ok', 'prod meta ok', 'mma first full', 'mma all issued', 'epi acc ready', 'epi done']
for name, run in cases.items():
    run(); torch.cuda.synchronize()
    run()
    buf = (C.c_ulonglong * (8 * 64))()
    fn(buf, 64)
    t = np.frombuffer(buf, dtype=np.uint64).reshape(64, 8).astype(np.int64)
    tiles = min(20, (n + 127) // 128 // 148)
    t0 = t[0, 0]
    print(f'== {name}: {tiles} tiles of CTA 0 (ns since its first tile start)')
    print('tile ' + ' '.join(f'{x:>18s}' for x in names))
    for j in range(tiles):
        print(f'{j:4d} ' + ' '.join(f'{int(t[j, s] - t0):18d}' for s in range(7)))
    d = np.diff(t[:tiles, 6])
    print('epilogue-done to epilogue-done per tile (ns): mean', d.mean(), ' | mma span (first full -> issued):', (t[:tiles, 4] - t[:tiles, 3]).mean(),
          '| epilogue span:', (t[:tiles, 6] - t[:tiles, 5]).mean(), '| prod wait tmem_empty:', (t[:tiles, 1] - t[:tiles, 0]).mean(),
          '| meta:', (t[:tiles, 2] - t[:tiles, 1]).mean(), '| meta ok -> first full:', (t[:tiles, 3] - t[:tiles, 2]).mean())
