"""Inputs of the whole-codec golden vectors (tests/golden/int_codec_golden.json): small topologies that walk every
block type of lossl_coord_int (OneScalePredictor with/without upsample, recurrent block with the 1-channel init
conv, OneScaleMultiStepPredictor with 2..4 steps, the wide multi-step variant, skipped top scales)."""
import numpy as np

from fastpcc_b200 import synth

CASES = [
    dict(name='c16_fea4', seed=1, n=3000, bits=9,
         cfg=dict(channels=16, max_stride_wo_recurrent=16, max_stride=64, fea_stride=4)),
    dict(name='c32_fea16', seed=2, n=2500, bits=9,
         cfg=dict(channels=32, max_stride_wo_recurrent=32, max_stride=128, fea_stride=16)),
    dict(name='c16_fea16_more_ch', seed=3, n=2000, bits=9,
         cfg=dict(channels=16, max_stride_wo_recurrent=32, max_stride=128, fea_stride=16,
                  use_more_ch_for_multi_step_pred=True)),
    dict(name='c64_skip_top', seed=4, n=3000, bits=9,
         cfg=dict(channels=64, max_stride_wo_recurrent=32, max_stride=128, fea_stride=16, skip_top_scales_num=1)),
    dict(name='c16_lidar', seed=1000, n=0, bits=16,
         cfg=dict(channels=16, max_stride_wo_recurrent=2048, max_stride=8192, fea_stride=16)),
    # the benchmarked width: channels >= 256 is the only configuration that takes the SparseConvPReLUIn8W8Out32 embed
    # branch of the 3- and 4-step predictors (reference model.py:130), and the tcgen05 kernels run whole 256-column tiles
    dict(name='c256_fea16', seed=5, n=2500, bits=9,
         cfg=dict(channels=256, max_stride_wo_recurrent=32, max_stride=128, fea_stride=16)),
    # BASELINE configs[1]/[2] topology (model_config.py:10-14 defaults / kitti_ford_ch128.yaml) on a 1:16 LiDAR frame
    dict(name='c128_lidar', seed=1001, n=0, bits=16,
         cfg=dict(channels=128, max_stride_wo_recurrent=2048, max_stride=8192, fea_stride=16)),
    dict(name='c256_lidar', seed=1002, n=0, bits=16,
         cfg=dict(channels=256, max_stride_wo_recurrent=2048, max_stride=8192, fea_stride=16)),
]


def case_cloud(case):
    """Unsorted (shuffled) unique voxels with a non-zero origin, int32 [N,3]."""
    if case['name'].endswith('lidar'):
        xyz = synth.lidar_frame(case['seed'])[::16]
    else:
        xyz = synth.surface_cloud(case['seed'], bits=case['bits'], n_target=case['n']) + np.array([5, 7, 11], np.int32)
    return np.ascontiguousarray(xyz[np.random.default_rng(case['seed']).permutation(xyz.shape[0])])
