"""Micro-benchmark of the fp16/bf16 fused sparse conv on a dense object surface (BASELINE configs[3]-like:
~800k occupied voxels of a 10-bit surface, C=128, 3x3x3).  usage: python tools/bench_conv_f16.py [n=800000] [channels=128]"""
import sys
import os.path as osp
import numpy as np
import torch
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from fastpcc_b200 import ops, synth  # noqa: E402

n_target = int(sys.argv[1]) if len(sys.argv) > 1 else 800000
ch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
C = torch.from_numpy(synth.with_batch(synth.surface_cloud(0, bits=10, n_target=n_target), 0)).cuda()
n = C.shape[0]
keys, vals = ops.hash_build(C)
table = ops.kmap_lookup(keys, vals, C, (3, 3, 3), (1, 1, 1))
pairs = int(torch.count_nonzero(table).item())


def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps


g = torch.Generator(device='cuda').manual_seed(0)
for dt in (torch.float16, torch.bfloat16):
    f = torch.randn((n, ch), generator=g, device='cuda').to(dt)
    w = (torch.randn((27, ch, ch), generator=g, device='cuda') / (27 * ch) ** 0.5).to(dt)
    b = torch.randn(ch, generator=g, device='cuda')
    ms = t(lambda: ops.spconv_f16(f, w, table, bias=b, act=ops.ACT_RELU))
    alg = 2.0 * pairs * ch * ch
    mma = 2.0 * ((n + 127) // 128 * 128) * 27 * ch * ch
    gb = (pairs * ch * 2 + n * ch * 2 + 27 * n * 4) / 1e9
    print(f'{dt}: surface n={n} C={ch} pairs/pt={pairs / n:.2f}: {ms:.3f} ms  algorithmic {alg / ms / 1e9:.1f} TFLOP/s  '
          f'executed {mma / ms / 1e9:.1f} TFLOP/s  gathered bytes {gb / ms * 1e3:.0f} GB/s')
    tp, perm = ops.group_rows(table)  # rows grouped by neighbour pattern: what the layer API does above GROUP_ROWS_MIN rows
    ms = t(lambda: ops.spconv_f16(f, w, tp, bias=b, act=ops.ACT_RELU, row_perm=perm))
    print(f'{dt}: same conv on grouped rows: {ms:.3f} ms  algorithmic {alg / ms / 1e9:.1f} TFLOP/s')
    w1 = (torch.randn((ch, ch), generator=g, device='cuda') / ch ** 0.5).to(dt)
    ms = t(lambda: ops.linear_f16(f, w1, bias=b, act=ops.ACT_RELU))
    print(f'{dt}: linear n={n} {ch}->{ch}: {ms:.3f} ms  {2.0 * n * ch * ch / ms / 1e9:.1f} TFLOP/s  {(2 * n * ch * 2) / ms / 1e6:.0f} GB/s')
