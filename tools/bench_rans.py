"""Micro-benchmark of the GPU range coder kernels (one or many streams); also the ncu target for them.
usage: python tools/bench_rans.py [n_symbols] [n_streams]"""
import sys
import os.path as osp
import numpy as np
import torch
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from fastpcc_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S = 255
g = torch.Generator(device='cuda').manual_seed(0)
logits = (torch.randn((n * ns, S), generator=g, device='cuda') * 3 * (1 << 23)).to(torch.int32)
cdf = ops.quantize_cdf(logits)
sym = torch.randint(0, S, (n * ns,), generator=g, device='cuda', dtype=torch.int32)
ranges = ops.cdf_symbol_ranges(logits, sym)
off = torch.arange(0, (ns + 1) * n, n, dtype=torch.int64, device='cuda')
cap = 2 * n + 64


def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps, r


ms, (out, out_len) = t(lambda: ops.rans_encode(ranges, off, cap))
print(f'encode: {ns} streams x {n} symbols: {ms:.3f} ms  {1e6 * ms / n:.1f} ns/symbol/stream')
lens = out_len.cpu()
blob = torch.cat([out[b, cap - int(lens[b]):] for b in range(ns)])
boff = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(lens.to(torch.int64), 0)[:-1]]).cuda()


def dec():
    d = ops.RansDecodeStreams(blob, boff, lens.cuda())
    return d.decode(cdf, S, off, n * ns), d


ms, (got, d) = t(dec)
print(f'decode: {ns} streams x {n} symbols: {ms:.3f} ms  {1e6 * ms / n:.1f} ns/symbol/stream  ok={bool((got == sym).all())} err={d.error()}')
