"""The multi-threaded CPU backend used by the timed CPU arm of bench.py (oracle/int_ops_fast.py) gives exactly the
integers of the numpy restatement (oracle/int_ops.py) operator by operator, and the reference's golden bitstreams
through the whole codec -- with the range coder swapped for the reference's own compiled C++ (oracle/_ref) as the
arm uses it."""
import hashlib
import json
import os.path as osp

import numpy as np
import pytest

from oracle import int_ops as K, int_ops_fast as F


def test_requant_family_equals_numpy():
    rng = np.random.default_rng(0)
    x = rng.integers(-(1 << 31), (1 << 31) - 1, (257, 48), endpoint=True).astype(np.int32)
    x[0, :6] = [-(1 << 31), (1 << 31) - 1, 0, -1, 1, -3]
    for dt in (np.int8, np.int16, np.int32):
        for shift in (0, 1, 9, 31, 37, 48, 62):
            for zpv in (0, -1, 77, -(1 << 40), 5 << min(shift, 57)):
                mul = rng.integers(0, (1 << 31) - 1, 48, endpoint=True).astype(np.uint32)  # products stay inside int64 (no wrap)
                mul[::9] = 0
                zp = np.array([zpv], np.int64)
                bias = rng.integers(-(1 << 29), 1 << 29, 48).astype(np.int32)
                for b in (None, bias):
                    for sl in (None, np.array([int(0.2 * (1 << 25))], np.int32), np.array([-(1 << 23)], np.int32), np.array([3 << 25], np.int32)):
                        xi = x >> 3
                        assert (F.requant(xi, mul, zp, shift, dt, bias=b, slope=sl) == K.requant(xi, mul, zp, shift, dt, bias=b, slope=sl)).all()
    for sv in (0, 1 << 25, int(0.3 * (1 << 25)), -(1 << 24), 5 << 25):
        s = np.array([sv], np.int32)
        assert (F.prelu(x, s) == K.prelu(x, s)).all()


def test_gemm_conv_softmax_equal_numpy():
    rng = np.random.default_rng(1)
    a = rng.integers(-128, 128, (300, 264)).astype(np.int8)
    w = rng.integers(-127, 128, (255, 264)).astype(np.int8)
    bias = rng.integers(-100000, 100000, 255).astype(np.int32)
    full = rng.integers(-1000, 1000, (300, 255)).astype(np.int32)
    for c in (None, bias, full):
        assert (F.gemm_int8(a, w, c) == K.gemm_int8(a, w, c)).all()
    big = rng.integers(-128, 128, (40, 1100)).astype(np.int8)  # beyond fp32 exactness: the float64 branch
    wb = rng.integers(-127, 128, (24, 1100)).astype(np.int8)
    assert (F.gemm_int8(big, wb) == K.gemm_int8(big, wb)).all()
    pts = np.unique(rng.integers(0, 40, (2500, 3)), axis=0).astype(np.int32)
    C = np.concatenate([np.zeros((pts.shape[0], 1), np.int32), pts], 1)
    f = rng.integers(-128, 128, (C.shape[0], 32)).astype(np.int8)
    wk = rng.integers(-127, 128, (27, 48, 32)).astype(np.int8)
    for same in (True, False):
        got, _ = F.sparse_conv_in8w8out32(f, wk, C, C, (3, 3, 3), (1, 1, 1), None, None, same)
        want, _ = K.sparse_conv_in8w8out32(f, wk, C, C, (3, 3, 3), (1, 1, 1), None, None, same)
        assert (got == want).all()
    zc = rng.integers(-500, 500, (27, 48)).astype(np.int32)
    got, _ = F.sparse_conv_in8w8out32(f, wk, C, C, (3, 3, 3), (1, 1, 1), None, zc, True)
    want, _ = K.sparse_conv_in8w8out32(f, wk, C, C, (3, 3, 3), (1, 1, 1), None, zc, True)
    assert (got == want).all()
    logits = rng.integers(-(1 << 27), 1 << 27, (500, 255)).astype(np.int32)
    logits[0] = 0
    logits[1] = -(1 << 31)
    assert (F.softmax_int32(logits >> 7) == K.softmax_int32(logits >> 7)).all()
    assert (F.batch_quantize_pmf(logits) == K.batch_quantize_pmf(logits)).all()


@pytest.mark.parametrize('name', ['c32_fea16', 'c16_lidar', 'c256_fea16'])
def test_whole_codec_on_the_fast_backend_reproduces_the_reference_golden(name):
    from fastpcc_b200 import synth
    from oracle import build_ref, lossl_coord_int as M
    from tests.golden.int_codec_cases import CASES, case_cloud
    gold = {g['name']: g for g in json.load(open(osp.join(osp.dirname(__file__), 'golden', 'int_codec_golden.json')))['cases']}[name]
    case = {c['name']: c for c in CASES}[name]
    ref = build_ref.load_ref('simple_rans_ext_cpp')
    saved = (M.K, M.RansEncoder, M.RansDecoder)
    try:
        M.K = F
        if ref is not None:  # the reference's own compiled coder, as in the timed arm
            M.RansEncoder, M.RansDecoder = ref.RansEncoder, ref.RansDecoder
        cfg = case['cfg']
        sd = synth.make_lossl_int_state_dict(seed=7, **{k: v for k, v in cfg.items() if k != 'skip_top_scales_num'})
        o = M.Model(sd, **cfg)
        data = o.compress(synth.with_batch(case_cloud(case)))
        assert hashlib.sha256(data).hexdigest() == gold['bitstream_sha256']
        rec = o.decompress(data)
        assert hashlib.sha256(np.ascontiguousarray(rec.astype('<i4')).tobytes()).hexdigest() == gold['decoded_sha256']
    finally:
        M.K, M.RansEncoder, M.RansDecoder = saved
