"""fpcc_nn_search / metrics.pc_error against the cKDTree oracle: exact squared distances, identical PSNR."""
import numpy as np
import pytest
import torch

from fastpcc_b200 import synth
from oracle import metrics as OM

pytestmark = pytest.mark.gpu


def _brute(q, r):
    d = ((q[:, None, :].astype(np.int64) - r[None].astype(np.int64)) ** 2).sum(2)
    return d.min(1), d.argmin(1)


@pytest.mark.parametrize('bits,nq,nr', [(6, 700, 1500), (10, 5000, 3000), (16, 3000, 5000), (20, 1030, 1025)])
def test_nn_search_exact(bits, nq, nr):
    from fastpcc_b200 import metrics
    rng = np.random.default_rng(bits)
    q = rng.integers(0, 1 << bits, (nq, 3)).astype(np.int32)
    r = rng.integers(0, 1 << bits, (nr, 3)).astype(np.int32)
    d2, idx = metrics.nn_search(torch.from_numpy(q).cuda(), torch.from_numpy(synth.with_batch(r)).cuda())
    want_d, want_i = _brute(q, r)
    assert (d2.cpu().numpy() == want_d).all()
    assert (idx.cpu().numpy() == want_i).all()  # lowest index among ties, as argmin


def test_pc_error_matches_oracle_and_identity_is_lossless():
    from fastpcc_b200 import metrics
    org = synth.surface_cloud(5, bits=10, n_target=30000)
    rng = np.random.default_rng(1)
    rec = np.unique(np.clip(org + rng.integers(-2, 3, org.shape), 0, 1023).astype(np.int32)[rng.random(org.shape[0]) < 0.8], axis=0)
    got = metrics.pc_error(torch.from_numpy(org).cuda(), torch.from_numpy(rec).cuda(), 1024, hausdorff=True)
    want = OM.pc_error(org, rec, 1024)
    for k, v in want.items():
        assert got[k] == pytest.approx(v, rel=1e-12), k
        assert round(got[k], 3) == round(v, 3)  # "identical D1-PSNR to 3 decimals"
    assert got['org points num'] == org.shape[0] and got['h.        (p2point)'] >= got['mseF      (p2point)']
    same = metrics.pc_error(torch.from_numpy(org).cuda(), torch.from_numpy(org).cuda(), 1024)
    assert same['mseF      (p2point)'] == 0 and same['mseF,PSNR (p2point)'] == float('inf')
    assert metrics.bpp(b'x' * 1000, 4000) == 2.0


def test_point_to_plane_uses_the_original_normals():
    from fastpcc_b200 import metrics
    # a z = 5 plane reconstructed at z = 7 with in-plane jitter: p2point > p2plane = 4 exactly
    g = np.stack(np.meshgrid(np.arange(0, 64, 2), np.arange(0, 64, 2), indexing='ij'), -1).reshape(-1, 2)
    org = np.concatenate([g, np.full((g.shape[0], 1), 5)], 1).astype(np.int32)
    rec = np.concatenate([g + 1, np.full((g.shape[0], 1), 7)], 1).astype(np.int32)
    nrm = torch.tensor([[0.0, 0.0, 1.0]]).repeat(org.shape[0], 1).cuda()
    out = metrics.pc_error(torch.from_numpy(org).cuda(), torch.from_numpy(rec).cuda(), 64, org_normals=nrm)
    assert out['mse1      (p2plane)'] == 4.0 and out['mse2      (p2plane)'] == 4.0
    assert out['mse1      (p2point)'] == 6.0
