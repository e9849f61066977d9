"""Micro-benchmark of the layer-to-layer requant (Q8.23 int32 -> scaled int8, optional fused PReLU): GB/s of
algorithmic traffic (5 B/elem).  usage: PYTHONPATH=. python tools/bench_requant.py [rows] [channels] [prelu=0|1]"""
import sys

import torch

from fastpcc_b200 import ops


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 3_250_000
    ch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    prelu = len(sys.argv) > 3 and sys.argv[3] == '1'
    x = torch.randint(-2 ** 30, 2 ** 30, (rows, ch), dtype=torch.int32, device='cuda')
    mul = torch.tensor([12345], dtype=torch.int32, device='cuda').view(torch.uint32)
    zp = torch.zeros(1, dtype=torch.int64, device='cuda')
    slope = torch.tensor([1 << 23], dtype=torch.int32, device='cuda') if prelu else None
    ep = ops.make_epilogue(mul, zp, 23 + 7, ops.OUT_I8, slope=slope)
    out = torch.empty((rows, ch), dtype=torch.int8, device='cuda')
    for _ in range(3):
        ops.requant(x, ep, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    n = 10
    for _ in range(n):
        ops.requant(x, ep, out=out)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / n
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n):
        y = x.to(torch.int8)
    t1.record()
    t1.synchronize()
    ms_t = t0.elapsed_time(t1) / n
    print(f'prelu={int(prelu)} rows={rows} ch={ch}: {ms:.3f} ms, {rows * ch * 5 / ms / 1e6:.0f} GB/s '
          f'(torch int32->int8 cast of the same tensor: {ms_t:.3f} ms, {rows * ch * 5 / ms_t / 1e6:.0f} GB/s)')


if __name__ == '__main__':
    main()
