"""Micro-benchmark of the layer-to-layer requant (Q8.23 int32 -> scaled int8): GB/s of algorithmic traffic (5 B/elem).
usage: python tests/bench_requant.py [rows] [channels]   (FPCC_REQUANT_VARIANT selects the kernel)"""
import os
import sys

import torch

from fastpcc_b200 import ops


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    ch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    x = torch.randint(-2 ** 30, 2 ** 30, (rows, ch), dtype=torch.int32, device='cuda')
    mul = torch.tensor([12345], dtype=torch.int32, device='cuda').view(torch.uint32)
    zp = torch.zeros(1, dtype=torch.int64, device='cuda')
    ep = ops.make_epilogue(mul, zp, 23 + 7, ops.OUT_I8)
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
    print(f'variant={os.environ.get("FPCC_REQUANT_VARIANT", "0")} rows={rows} ch={ch}: {ms:.3f} ms, '
          f'{rows * ch * 5 / ms / 1e6:.0f} GB/s, checksum {int(out.view(torch.uint8).sum(dtype=torch.int64))}')


if __name__ == '__main__':
    main()
