"""Micro-benchmark of the fused sparse-conv / linear kernels on one LiDAR pyramid level (also the ncu target).
usage: python tools/bench_conv.py [stride_log2=4] [channels=256] [frames=1]"""
import sys
import os.path as osp
import numpy as np
import torch
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from fastpcc_b200 import ops, synth  # noqa: E402

lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 1
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
mul = torch.from_numpy(rng.integers(1 << 18, 1 << 22, ch).astype(np.int64)).to(torch.uint32).cuda()
zp = torch.zeros(1, dtype=torch.int64, device='cuda')
slope = torch.tensor([1 << 23], dtype=torch.int32, device='cuda')
keys, vals = ops.hash_build(C)
table = ops.kmap_lookup(keys, vals, C, (3, 3, 3), (1, 1, 1))
pairs = int(torch.count_nonzero(table).item())
ep = ops.make_epilogue(mul, zp, 24, ops.OUT_I8, bias=bias, slope=slope)
ep32 = ops.make_epilogue(mul, zp, 6, ops.OUT_I32, bias=bias, slope=slope)  # overflows int32: every chunk falls back to the 64-bit epilogue
# the regime of a PTQ-converted model's int32 (Q8.23) producing layers: multipliers near 2^30, shift 37 (high-word fast path)
mul_r = torch.from_numpy(rng.integers(1 << 29, 1 << 30, ch).astype(np.int64)).to(torch.uint32).cuda()
ep32r = ops.make_epilogue(mul_r, zp, 37, ops.OUT_I32, bias=bias, slope=slope)


def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps


alg = 2.0 * pairs * ch * ch
mma = 2.0 * ((n + 127) // 128 * 128) * 27 * ch * ch
w2 = torch.from_numpy(rng.integers(-127, 128, (ch, ch)).astype(np.int8)).cuda()
occ = torch.from_numpy(rng.integers(1, 256, n).astype(np.uint8)).cuda()
rb_table = torch.from_numpy(rng.integers(-100000, 100000, (256, ch)).astype(np.int32)).cuda()
res = torch.from_numpy(rng.integers(-(1 << 28), 1 << 28, (n, ch)).astype(np.int32)).cuda()
# regimes of a PTQ-converted model (fastpcc_b200/synth.py): int8-producing layers shift 28..39, int32 (Q8.23) producing
# layers shift - 23 = 5..16, multipliers just below 2^(32 - guard bits)
mul_hi = torch.from_numpy(rng.integers(1 << 29, 1 << 30, ch).astype(np.int64)).to(torch.uint32).cuda()
mul_lo = torch.from_numpy(rng.integers(1 << 20, 1 << 21, ch).astype(np.int64)).to(torch.uint32).cuda()
cases = {
    'i8  shift 24 prelu         ': ops.make_epilogue(mul, zp, 24, ops.OUT_I8, bias=bias, slope=slope),
    'i8  shift 38 prelu         ': ops.make_epilogue(mul_hi, zp, 38, ops.OUT_I8, bias=bias, slope=slope),
    'i8  shift 38               ': ops.make_epilogue(mul_hi, zp, 38, ops.OUT_I8, bias=bias),
    'i8  shift 38 prelu rowbias ': ops.make_epilogue(mul_hi, zp, 38, ops.OUT_I8, bias=bias, slope=slope, row_bias=(rb_table, occ, 100000)),
    'i32 shift 12               ': ops.make_epilogue(mul_lo, zp, 12, ops.OUT_I32, bias=bias),
    'i32 shift 12 prelu rowbias ': ops.make_epilogue(mul_lo, zp, 12, ops.OUT_I32, bias=bias, slope=slope, row_bias=(rb_table, occ, 100000)),
    'i32 shift 37 prelu         ': ops.make_epilogue(mul_hi, zp, 37, ops.OUT_I32, bias=bias, slope=slope),
    'i32 shift 6 (saturating)   ': ops.make_epilogue(mul, zp, 6, ops.OUT_I32, bias=bias, slope=slope),
}
ms = t(lambda: ops.spconv(f, w, table, ep))
print(f'conv  stride 2^{lvl} n={n} C={ch} pairs/pt={pairs / n:.2f}: {ms:.3f} ms  algorithmic {alg / ms / 1e9:.1f} TOP/s  executed {mma / ms / 1e9:.1f} TOP/s')
for name, e in cases.items():
    ms = t(lambda: ops.linear(f, w2, e))
    esz = 1 if e.out_type == ops.OUT_I8 else 4
    print(f'linear {name} n={n} {ch}->{ch}: {ms:.3f} ms  {2.0 * n * ch * ch / ms / 1e9:7.1f} TOP/s  {n * ch * (1 + esz) / ms / 1e6:6.0f} GB/s')
tp0, perm0 = ops.group_rows(table)
epr = ops.make_epilogue(mul_lo, zp, 12, ops.OUT_I32, bias=bias, residual=res, post_slope=slope)
ms = t(lambda: ops.spconv(f, w, tp0, epr, row_perm=perm0))
print(f'conv grouped, i32 out + residual + post PReLU (ResBlock conv2): {ms:.3f} ms  algorithmic {alg / ms / 1e9:.1f} TOP/s')
eph = ops.make_epilogue(mul_hi, zp, 38, ops.OUT_I8, bias=bias, slope=slope)
ms = t(lambda: ops.spconv(f, w, tp0, eph, row_perm=perm0))
print(f'conv grouped, i8 out shift 38 prelu (ResBlock conv_prelu): {ms:.3f} ms  algorithmic {alg / ms / 1e9:.1f} TOP/s')
ms = t(lambda: ops.group_rows(table))
tp, perm = ops.group_rows(table)
print(f'group_rows n={n}: {ms:.3f} ms')
ms = t(lambda: ops.spconv(f, w, tp, ep, row_perm=perm))
print(f'conv on grouped rows: {ms:.3f} ms  algorithmic {alg / ms / 1e9:.1f} TOP/s   equal={bool((ops.spconv(f, w, tp, ep, row_perm=perm) == ops.spconv(f, w, table, ep)).all())}')
ms = t(lambda: ops.kmap_lookup(keys, vals, C, (3, 3, 3), (1, 1, 1)))
print(f'kmap lookup n={n}: {ms:.3f} ms  {(16 * n + 8 * 27 * n + 4 * 27 * n) / ms / 1e6:.1f} GB/s (algorithmic bytes)')
ms = t(lambda: ops.hash_build(C))
print(f'hash build n={n}: {ms:.3f} ms')
