"""End-to-end parity of the lossl_coord_int codec on the GPU against the CPU oracle: byte-identical
bitstreams, identical per-level symbol ranges, lossless reconstruction in the reference's point order."""
import numpy as np
import pytest
import torch

from fastpcc_b200 import synth
from oracle.lossl_coord_int import Model as OracleModel

pytestmark = pytest.mark.gpu

CFGS = [
    dict(channels=16, max_stride_wo_recurrent=16, max_stride=64, fea_stride=4),
    dict(channels=32, max_stride_wo_recurrent=32, max_stride=128, fea_stride=16),
    dict(channels=16, max_stride_wo_recurrent=32, max_stride=128, fea_stride=16, use_more_ch_for_multi_step_pred=True),
    dict(channels=64, max_stride_wo_recurrent=32, max_stride=128, fea_stride=16, skip_top_scales_num=1),
]


def _make(cfg):
    from fastpcc_b200.lossl_coord_int import Config, Model
    sd = synth.make_lossl_int_state_dict(seed=7, **{k: v for k, v in cfg.items() if k != 'skip_top_scales_num'})
    m = Model(Config(**cfg), device='cuda').load_numpy_state_dict(sd).cuda()
    return m, OracleModel(sd, **cfg)


@pytest.mark.parametrize('cfg', CFGS)
def test_bitstream_is_byte_identical_and_lossless(cfg):
    m, o = _make(cfg)
    xyz = synth.surface_cloud(1, bits=9, n_target=3000) + np.array([5, 7, 11], np.int32)
    xyz = xyz[np.random.default_rng(0).permutation(xyz.shape[0])]
    want = o.compress(synth.with_batch(xyz))
    got = m.compress(torch.from_numpy(synth.with_batch(xyz)).cuda())
    assert got[:8] == want[:8]
    assert got == want
    rec = m.decompress(got).cpu().numpy()
    assert (rec == o.decompress(want)).all()
    assert (np.unique(rec, axis=0) == np.unique(xyz, axis=0)).all()


def test_lidar_frame_roundtrip_c128():
    cfg = dict(channels=128, max_stride_wo_recurrent=2048, max_stride=8192, fea_stride=16)
    from fastpcc_b200.lossl_coord_int import Config, Model
    sd = synth.make_lossl_int_state_dict(seed=7, **cfg)
    m = Model(Config(**cfg), device='cuda').load_numpy_state_dict(sd).cuda()
    xyz = synth.lidar_frame(1000)[::4]
    data = m.compress(torch.from_numpy(synth.with_batch(xyz)).cuda())
    rec = m.decompress(data).cpu().numpy()
    assert rec.shape == xyz.shape and (np.unique(rec, axis=0) == np.unique(xyz, axis=0)).all()
