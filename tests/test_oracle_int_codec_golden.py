"""Pins the numpy restatement of the whole codec (oracle/lossl_coord_int.py) to the reference's own Python:
tests/golden/int_codec_golden.json was minted by tests/golden/make_int_codec_golden.py, which runs the unmodified
models/convolutional/lossl_coord_int/model.py + lib/int_sparse_conv/cuda_ops.py + the reference's compiled range
coder on the CPU (only the CUDA extension's 20 entry points are stood in for).  Bit-exact: bytes and decoded order."""
import hashlib
import json
import os.path as osp

import numpy as np
import pytest

from fastpcc_b200 import synth
from oracle.lossl_coord_int import Model
from tests.golden.int_codec_cases import CASES, case_cloud

GOLDEN = json.load(open(osp.join(osp.dirname(__file__), 'golden', 'int_codec_golden.json')))['cases']


def test_golden_covers_every_case():
    assert [g['name'] for g in GOLDEN] == [c['name'] for c in CASES]


@pytest.mark.parametrize('case,gold', list(zip(CASES, GOLDEN)), ids=[c['name'] for c in CASES])
def test_oracle_codec_matches_reference_python(case, gold):
    cfg = case['cfg']
    sd = synth.make_lossl_int_state_dict(seed=7, **{k: v for k, v in cfg.items() if k != 'skip_top_scales_num'})
    o = Model(sd, **cfg)
    xyz = case_cloud(case)
    assert xyz.shape[0] == gold['n_points']
    data = o.compress(synth.with_batch(xyz))
    assert len(data) == gold['n_bytes']
    assert hashlib.sha256(data).hexdigest() == gold['bitstream_sha256']
    rec = o.decompress(data)
    assert hashlib.sha256(np.ascontiguousarray(rec.astype('<i4')).tobytes()).hexdigest() == gold['decoded_sha256']
