"""Pins the numpy restatement of lib/int_sparse_conv to the reference source: LUT hash, rounding and
saturation rules by hand-computed cases, kernel-offset enumeration, and whole-codec round trips."""
import hashlib

import numpy as np
import pytest

from fastpcc_b200 import synth
from oracle import int_ops as K
from oracle.lossl_coord_int import Model

# SHA-256 of the 6145-entry int32 literal at lib/int_sparse_conv/src/softmax.cu:13-22 (little endian),
# computed from /root/reference when this test was written.
LUT_SHA256 = '72f3293e649b586848bccaed67dd0fa693f222e80b12c9b5f6c82d1ff9dafd88'


def test_exp_lut_matches_reference_literal():
    lut = K.exp_lut()
    assert lut.shape == (6145,) and lut[0] == 65536 and lut[-1] == 0
    assert hashlib.sha256(lut.astype('<i4').tobytes()).hexdigest() == LUT_SHA256


def test_round_half_away():
    x = np.array([5, -5, 3, -3, 4, -4, 0, 1, -1], dtype=np.int64)
    assert K.rha_shift(x, 1).tolist() == [3, -3, 2, -2, 2, -2, 0, 1, -1]
    assert K.rha_shift(x, 0).tolist() == x.tolist()
    assert K.rha_shift(np.array([6, -6, 2, -2]), 2).tolist() == [2, -2, 1, -1]


def test_requant_variants_by_hand():
    inp = np.array([[100, -100], [2 ** 31 - 1, -2 ** 31]], dtype=np.int32)
    mul = np.array([3, 2 ** 32 - 1], dtype=np.uint32)
    zp = np.array([7], dtype=np.int64)
    out = K.requant(inp, mul, zp, 2, np.int32)
    # (100*3+7)=307 -> (307+2)>>2 = 77 ; (-100*(2^32-1)+7) -> -(429496729493+2 >>2)
    assert out[0, 0] == 77
    assert out[0, 1] == np.clip(-((100 * (2 ** 32 - 1) - 7 + 2) >> 2), -2 ** 31, 2 ** 31 - 1)
    assert out[1, 0] == (((2 ** 31 - 1) * 3 + 7 + 2) >> 2)
    assert out[1, 1] == -2 ** 31
    o8 = K.requant(inp, mul, zp, 0, np.int8)
    assert o8.tolist() == [[127, -128], [127, -128]]
    # bias then Q6.25 PReLU on negatives only (slope 0.25): v=-100+(-28) = -128 -> -32
    sl = np.array([1 << 23], dtype=np.int32)
    o = K.requant(np.array([[-100, 50]], dtype=np.int32), np.array([1, 1], dtype=np.uint32), np.zeros(1, np.int64),
                  0, np.int32, bias=np.array([-28, 3], dtype=np.int32), slope=sl)
    assert o.tolist() == [[-32, 53]]
    # prelu rounding half away: -2 * 0.25 = -0.5 -> -1 ; -1*0.25 = -0.25 -> 0
    assert K.prelu(np.array([[-2, -1, 2, -6]], dtype=np.int32), sl).tolist() == [[-1, 0, 2, -2]]


def test_kernel_offsets_enumeration():
    o3 = K.kernel_offsets((3, 3, 3))
    assert o3[0].tolist() == [-1, -1, -1] and o3[1].tolist() == [0, -1, -1] and o3[13].tolist() == [0, 0, 0]
    assert o3[26].tolist() == [1, 1, 1] and o3[3].tolist() == [-1, 0, -1]
    o2 = K.kernel_offsets((2, 2, 2))
    assert o2.tolist() == [[(k >> 2) & 1, (k >> 1) & 1, k & 1] for k in range(8)]  # z fastest, == unfold_kernel
    o4 = K.kernel_offsets((4, 4, 4))
    assert o4.min() == -1 and o4.max() == 2 and o4[1].tolist() == [-1, -1, 0]  # the (ks-1)//2 shift of the .cuh


def test_softmax_rows_sum_to_one_and_cdf_is_monotone():
    rng = np.random.default_rng(0)
    logits = (rng.normal(0, 3, (64, 255)) * (1 << 23)).astype(np.int32)
    p = K.softmax_int32(logits >> 7)
    s = p.astype(np.float64).sum(1) / 2 ** 32
    assert np.abs(s - 1).max() < 1e-3
    cdf = K.batch_quantize_pmf(logits).astype(np.int64)
    assert (np.diff(cdf, axis=1) >= 1)[:, :-1].all() and (cdf[:, -1] == 65535).all() and (cdf[:, 0] >= 1).all()
    assert (cdf[:, -2] < 65535).all()


def test_sparse_conv_matches_bruteforce():
    rng = np.random.default_rng(1)
    pts = np.unique(rng.integers(0, 12, (300, 3)), axis=0).astype(np.int32)
    C = synth.with_batch(pts)
    f = rng.integers(-127, 128, (C.shape[0], 8)).astype(np.int8)
    w = rng.integers(-127, 128, (27, 5, 8)).astype(np.int8)
    out, maps = K.sparse_conv_in8w8out32(f, w, C, C, (3, 3, 3), (1, 1, 1), None, None, True)
    assert maps[13] == (None, None)
    lut = {tuple(p): i for i, p in enumerate(pts.tolist())}
    ref = np.zeros_like(out)
    for i, p in enumerate(pts.tolist()):
        for k, (dx, dy, dz) in enumerate(K.kernel_offsets((3, 3, 3)).tolist()):
            j = lut.get((p[0] + dx, p[1] + dy, p[2] + dz))
            if j is not None:
                ref[i] += w[k].astype(np.int32) @ f[j].astype(np.int32)
    assert (out == ref).all()


@pytest.mark.parametrize('cfg', [
    dict(channels=16, max_stride_wo_recurrent=16, max_stride=64, fea_stride=4),
    dict(channels=16, max_stride_wo_recurrent=32, max_stride=128, fea_stride=16),
    dict(channels=16, max_stride_wo_recurrent=32, max_stride=128, fea_stride=16, use_more_ch_for_multi_step_pred=True),
])
def test_codec_roundtrip_is_lossless(cfg):
    sd = synth.make_lossl_int_state_dict(seed=7, **cfg)
    xyz = synth.surface_cloud(1, bits=9, n_target=1500) + np.array([5, 7, 11], np.int32)
    xyz = xyz[np.random.default_rng(0).permutation(xyz.shape[0])]
    m = Model(sd, **cfg)
    data = m.compress(synth.with_batch(xyz))
    rec = m.decompress(data)
    assert int.from_bytes(data[:2], 'little') == xyz[:, 0].min()
    assert (np.unique(rec, axis=0) == np.unique(xyz, axis=0)).all() and rec.shape == xyz.shape
