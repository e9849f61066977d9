"""GPU front end (fpcc_voxelize_f32, fpcc_kd_split through fastpcc_b200/frontend.py) against the oracle and the
reference-minted kd partitions: bit-exact voxels, order and partitions."""
import hashlib
import json
import os.path as osp

import numpy as np
import pytest
import torch

from fastpcc_b200 import synth
from oracle import frontend as OF
from tests.golden.frontend_cases import KD_CASES, kd_cloud

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(osp.join(osp.dirname(__file__), 'golden', 'frontend_golden.json')))


def _h(a):
    return hashlib.sha256(np.ascontiguousarray(a.astype('<i4')).tobytes()).hexdigest()


@pytest.mark.parametrize('seed,ld', [(1003, 3), (1004, 4)])
def test_voxelize_matches_numpy_transform(seed, ld):
    from fastpcc_b200 import frontend
    pts = synth.lidar_points(seed)
    if ld == 4:  # KITTI .bin rows: x, y, z, reflectance
        pts = np.concatenate([pts, np.random.default_rng(0).random((pts.shape[0], 1), dtype=np.float32)], 1)
    want, org, inv = OF.voxelize(np.ascontiguousarray(pts[:, :3]))
    got, inv_t = frontend.voxelize(torch.from_numpy(pts).cuda(), batch=3)
    got = got.cpu().numpy()
    assert got.shape == (want.shape[0], 4) and (got[:, 0] == 3).all()
    assert (got[:, 1:] == want).all()
    inv_t = inv_t.cpu().numpy()
    assert (inv_t[:3] == org).all() and inv_t[3] == np.float32(inv)


def test_voxelize_edge_cases():
    from fastpcc_b200 import frontend
    # ties at .5 (round half to even), negative coordinates, duplicates, a single point
    pts = np.array([[-1.0, 0.0, 2.0], [-0.5, 0.5, 2.5], [0.5, 1.5, 3.5], [0.5, 1.5, 3.5], [1.5, 2.5, 4.5]], np.float32)
    want, org, _ = OF.voxelize(pts, resolution=401, span=400.0)  # scale 1.0
    got, _ = frontend.voxelize(torch.from_numpy(pts).cuda(), resolution=401, span=400.0)
    assert (got.cpu().numpy()[:, 1:] == want).all()
    one, _ = frontend.voxelize(torch.tensor([[3.0, -2.0, 7.5]]).cuda())
    assert one.cpu().tolist() == [[0, 0, 0, 0]]
    with pytest.raises(RuntimeError):  # does not fit the grid
        frontend.voxelize(torch.tensor([[0.0, 0.0, 0.0], [500.0, 0.0, 0.0]]).cuda(), resolution=1024, span=400.0)
    with pytest.raises(RuntimeError):
        frontend.voxelize(torch.zeros((4, 3)))  # CPU tensor: no fallback


@pytest.mark.parametrize('case', KD_CASES, ids=[c['name'] for c in KD_CASES])
def test_kd_partition_matches_reference_golden(case):
    from fastpcc_b200 import frontend
    xyz = kd_cloud(case)
    gold = GOLDEN[case['name']]
    parts = frontend.kd_tree_partition(torch.from_numpy(synth.with_batch(xyz)).cuda(), case['max_num'])
    assert [p.shape[0] for p in parts] == gold['sizes']
    assert [_h(p[:, 1:].cpu().numpy()) for p in parts] == gold['sha256']


def test_partitioned_cloud_codes_to_the_partition_container():
    """voxelize -> kd partition -> compress_partitions -> decompress_partitions returns every voxel once."""
    from fastpcc_b200 import frontend
    from fastpcc_b200.lossl_coord_int import Config, Model
    cfg = dict(channels=16, max_stride_wo_recurrent=32, max_stride=128, fea_stride=16)
    sd = synth.make_lossl_int_state_dict(seed=7, **cfg)
    m = Model(Config(**cfg), device='cuda').load_numpy_state_dict(sd).cuda()
    coords, _ = frontend.voxelize(torch.from_numpy(synth.lidar_points(1005)[::6]).cuda(), resolution=4096)
    parts = frontend.collate_partitions(coords, 6000)
    assert len(parts) > 2
    blob = m.compress_partitions(parts)
    rec = m.decompress_partitions(blob).cpu().numpy()
    assert (np.unique(rec, axis=0) == np.unique(coords[:, 1:].cpu().numpy(), axis=0)).all() and rec.shape[0] == coords.shape[0]
