"""oracle/frontend.py against the reference's own kd_tree_partition (golden minted from lib/data_utils.py)."""
import hashlib
import json
import os.path as osp

import numpy as np
import pytest

from oracle import frontend as F
from tests.golden.frontend_cases import KD_CASES, kd_cloud

GOLDEN = json.load(open(osp.join(osp.dirname(__file__), 'golden', 'frontend_golden.json')))


def _h(a):
    return hashlib.sha256(np.ascontiguousarray(a.astype('<i4')).tobytes()).hexdigest()


@pytest.mark.parametrize('case', KD_CASES, ids=[c['name'] for c in KD_CASES])
def test_kd_partition_oracle_matches_reference(case):
    xyz = kd_cloud(case)
    gold = GOLDEN[case['name']]
    parts = F.kd_tree_partition(xyz, case['max_num'])
    assert [len(p) for p in parts] == gold['sizes']
    assert [_h(p) for p in parts] == gold['sha256']


def test_voxelize_is_unique_sorted_and_idempotent_on_grid_points():
    from fastpcc_b200 import synth
    from oracle.lossl_coord_int import morton_xmajor
    pts = synth.lidar_points(1002)
    q, org, inv = F.voxelize(pts)
    assert q.dtype == np.int32 and q.min() == 0 and q.max() < 65536
    assert (np.diff(morton_xmajor(q)) > 0).all()
    assert np.allclose(org, pts.min(0)) and abs(inv - 400 / 65535) < 1e-12
    # the transform of synth.lidar_frame (the bench workload) is this function
    assert (np.unique(q, axis=0) == synth.lidar_frame(1002)).all()
