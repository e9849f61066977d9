"""Inputs of the float (lossy) codec fixtures: the reference's lossy_coord_v2 model at its baseline_r1 topology
(config/convolutional/lossy_coord_v2/baseline_r1.yaml) with seeded parameters, on small 10-bit surface clouds."""
import numpy as np

from fastpcc_b200 import synth

CASES = [
    dict(name='v2_r1_20k', seed=1, n=20000, bits=10, model_seed=0, bottleneck_scaler=32),
    dict(name='v2_r1_6k', seed=2, n=6000, bits=10, model_seed=3, bottleneck_scaler=32),
]


def case_cloud(case):
    xyz = synth.surface_cloud(case['seed'], bits=case['bits'], n_target=case['n']) + np.array([3, 0, 9], np.int32)
    return np.ascontiguousarray(xyz[np.random.default_rng(case['seed']).permutation(xyz.shape[0])])
