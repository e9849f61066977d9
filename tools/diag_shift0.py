"""Diagnostic for the open question in DESIGN 3.5 (not collected by pytest): shift 0 with a non-zero zero point and
non-saturating negatives, through the fused linear epilogue and the stand-alone requant; prints every mismatch with
got / want instead of asserting.  usage: python tools/diag_shift0.py"""
import os.path as osp
import sys

import numpy as np
import torch

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from fastpcc_b200 import ops  # noqa: E402
from oracle import int_ops as K  # noqa: E402  (test infrastructure)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def report(tag, got, want):
    bad = np.argwhere(got != want)
    print(f'{tag}: {"ok" if len(bad) == 0 else f"{len(bad)} mismatches"}')
    for i in bad[:6]:
        i = tuple(int(v) for v in i)
        print(f'    at {i}: got {int(got[i])} want {int(want[i])}')


def main():
    rng = np.random.default_rng(7)
    m, k, n = 300, 64, 96
    a = rng.integers(-128, 128, (m, k)).astype(np.int8)
    w = rng.integers(-127, 128, (n, k)).astype(np.int8)
    acc0 = K.gemm_int8(a, w, None)
    small = (acc0 >> 12).astype(np.int32)
    a_s, w_s = np.zeros((m, k), np.int8), np.zeros((n, k), np.int8)
    for out_t, np_t in ((ops.OUT_I8, np.int8), (ops.OUT_I16, np.int16), (ops.OUT_I32, np.int32)):
        for zpv in (0, 1, -1, 5):
            mul1, zp = np.ones(n, np.uint32), np.array([zpv], np.int64)
            bias_s = small[zpv % m].copy()
            want = K.requant(np.zeros((m, n), np.int32), mul1, zp, 0, np_t, bias=bias_s)
            got = ops.linear(dev(a_s), dev(w_s), ops.make_epilogue(dev(mul1), dev(zp), 0, out_t, bias=dev(bias_s))).cpu().numpy()
            report(f'linear shift 0 out{out_t} zp {zpv}', got, want)
    x = rng.integers(-(1 << 31), (1 << 31) - 1, (513, 64), endpoint=True).astype(np.int32)
    x[0, :8] = [-(1 << 31), (1 << 31) - 1, 0, -1, 1, -(1 << 31) + 1, 1 << 30, -(1 << 30)]
    x[1, :8] = [-10, -3, -2, -7, -100, 9, -41, -127]
    for shift in (0, 1, 7, 23, 31, 32, 36, 48, 62):
        for mulv in (0, 1, 3, 12345, (1 << 22) + 5, (1 << 30) + 99, (1 << 31) - 1, (1 << 31) + 7):
            for zpv in (0, -1, 5 << min(shift, 57), -(1 << 40), (1 << 60) - 1, -(3 << 58)):
                mul1, zp = np.array([mulv], np.uint32), np.array([zpv], np.int64)
                want = K.requant(x, np.full(64, mulv, np.uint32), zp, shift, np.int8)
                got = ops.requant(dev(x), ops.make_epilogue(dev(mul1), dev(zp), shift, ops.OUT_I8)).cpu().numpy()
                if (got != want).any():
                    report(f'requant shift {shift} mul {mulv} zp {zpv}', got, want)
    print('done')


if __name__ == '__main__':
    main()
