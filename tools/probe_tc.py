"""Stand-alone A/B probe of the tcgen05 kernels against the CUDA-core kernels (same C ABI, same inputs).
Run it under `timeout` before the GPU test-suite: a wrong descriptor shows up here as a diff pattern (or a
hang that the timeout ends) instead of taking the whole suite down.  Prints one line per case."""
import sys
import os.path as osp

import numpy as np
import torch

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from fastpcc_b200 import _lib, ops  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def describe(got, want, tag):
    bad = got != want
    nbad = int(bad.sum())
    if nbad == 0:
        print(f'OK   {tag}')
        return True
    rows = np.nonzero(bad.any(1))[0]
    cols = np.nonzero(bad.any(0))[0]
    print(f'FAIL {tag}: {nbad}/{bad.size} wrong; rows {rows[:8].tolist()}..({len(rows)}) cols {cols[:8].tolist()}..({len(cols)})')
    r, c = np.argwhere(bad)[0]
    print(f'     first: [{r},{c}] got {got[r, c]} want {want[r, c]}; row got {got[r, :6].tolist()} want {want[r, :6].tolist()}')
    return False


def run_linear(m, k, n, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.integers(-128, 128, (m, k)).astype(np.int8)
    w = rng.integers(-127, 128, (n, k)).astype(np.int8)
    mul = torch.ones(1, dtype=torch.int32).view(torch.uint32).cuda()
    zp = torch.zeros(1, dtype=torch.int64, device='cuda')
    res = []
    for mode in (0, 1):
        _lib.load().fpcc_set_tc_mode(mode)
        ep = ops.make_epilogue(mul, zp, 0, ops.OUT_I32)
        res.append(ops.linear(dev(a), dev(w), ep).cpu().numpy())
        torch.cuda.synchronize()
    want = a.astype(np.int64) @ w.astype(np.int64).T
    ok0 = describe(res[0], want.astype(np.int32), f'simt linear m={m} k={k} n={n}')
    ok1 = describe(res[1], want.astype(np.int32), f'tc   linear m={m} k={k} n={n}')
    return ok0 and ok1


def run_conv(n_pts, cin, cout, bits, seed=0):
    rng = np.random.default_rng(seed)
    pts = np.unique(rng.integers(0, 1 << bits, (n_pts, 3)), axis=0).astype(np.int32)
    C = np.concatenate([np.zeros((pts.shape[0], 1), np.int32), pts], 1)
    f = rng.integers(-128, 128, (C.shape[0], cin)).astype(np.int8)
    w = rng.integers(-127, 128, (27, cout, cin)).astype(np.int8)
    keys, vals = ops.hash_build(dev(C))
    table = ops.kmap_lookup(keys, vals, dev(C), (3, 3, 3), (1, 1, 1))
    pairs = int((table != 0).sum().item())
    res = []
    for mode in (0, 1):
        _lib.load().fpcc_set_tc_mode(mode)
        res.append(ops.spconv(dev(f), dev(w), table, ops.identity_epilogue(torch.device('cuda', 0))).cpu().numpy())
        torch.cuda.synchronize()
    return describe(res[1], res[0], f'tc vs simt conv n={C.shape[0]} pairs/pt={pairs / C.shape[0]:.1f} cin={cin} cout={cout}')


if __name__ == '__main__':
    assert torch.cuda.is_available()
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    ok = True
    for m, k, n in [(128, 32, 16), (128, 128, 32), (128, 128, 256), (100, 64, 48), (1000, 256, 256), (333, 512, 255),
                    (4096, 256, 2048), (257, 272, 64)]:
        ok &= run_linear(m, k, n)
    for n_pts, cin, cout, bits in [(2000, 32, 32, 5), (5000, 128, 128, 6), (20000, 256, 256, 7), (3000, 64, 255, 5)]:
        ok &= run_conv(n_pts, cin, cout, bits)
    print('PROBE', 'PASS' if ok else 'FAIL')
    sys.exit(0 if ok else 1)
