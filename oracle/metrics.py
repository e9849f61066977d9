"""CPU restatement of the geometry figures of MPEG `pc_error` (as parsed by lib/metrics/pc_error_wrapper.py:76-98).

TEST INFRASTRUCTURE ONLY.  Parity status: "parity unpinned" -- the pc_error binary and its source are not part of the
reference repository (scripts/script_config.py only holds a path), so the definition below is the published one:
nearest-neighbour squared distances in both directions, mse = mean, symmetric = max, PSNR = 10 log10(3 peak^2 / mse)
with peak = resolution - 1 (pc_error_wrapper.py:50).  Exact NN distances come from scipy's cKDTree."""
import math

import numpy as np
from scipy.spatial import cKDTree


def nn_d2(query: np.ndarray, ref: np.ndarray):
    d, i = cKDTree(ref.astype(np.float64)).query(query.astype(np.float64), k=1)
    diff = query.astype(np.int64) - ref.astype(np.int64)[i]
    return (diff * diff).sum(1), i


def pc_error(org: np.ndarray, rec: np.ndarray, resolution: float):
    peak = resolution - 1
    m1 = float(nn_d2(org, rec)[0].mean())
    m2 = float(nn_d2(rec, org)[0].mean())
    mf = max(m1, m2)
    ps = lambda m: float('inf') if m == 0 else 10 * math.log10(3 * peak * peak / m)  # noqa: E731
    return {'mse1      (p2point)': m1, 'mse1,PSNR (p2point)': ps(m1), 'mse2      (p2point)': m2,
            'mse2,PSNR (p2point)': ps(m2), 'mseF      (p2point)': mf, 'mseF,PSNR (p2point)': ps(mf)}
