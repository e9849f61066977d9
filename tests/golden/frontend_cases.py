"""Inputs of the front-end golden vectors (tests/golden/frontend_golden.json)."""
import numpy as np

from fastpcc_b200 import synth

KD_CASES = [
    dict(name='lidar_16bit_30k', max_num=30000),
    dict(name='surface_10bit_6k', max_num=6000),
    dict(name='plane_ties_500', max_num=500),
    dict(name='small_no_split', max_num=5000),
    dict(name='one_level', max_num=1500),
]


def kd_cloud(case):
    """int32 [N,3] voxels in the order the model receives them (Morton, x most significant)."""
    from oracle.lossl_coord_int import morton_xmajor
    name = case['name']
    if name.startswith('lidar'):
        xyz = synth.lidar_frame(1001)
    elif name.startswith('surface'):
        xyz = synth.surface_cloud(21, bits=10, n_target=40000)
    elif name.startswith('plane'):  # many equal coordinates on the split axis: exercises `<= split_value`
        rng = np.random.default_rng(3)
        xyz = np.unique(np.stack([rng.integers(0, 40, 4000) * 8, rng.integers(0, 6, 4000), rng.integers(0, 300, 4000)], 1)
                        .astype(np.int32), axis=0)
    elif name.startswith('small'):
        xyz = synth.surface_cloud(22, bits=8, n_target=3000)
    else:
        xyz = synth.surface_cloud(23, bits=9, n_target=2500)
    return np.ascontiguousarray(xyz[np.argsort(morton_xmajor(xyz), kind='stable')])
