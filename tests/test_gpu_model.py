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


def test_batched_frames_equal_single_frame_streams():
    """Frames coded together (one launch per kernel, one rANS stream per frame) give exactly the bytes of
    coding each frame alone -- ragged sizes, different offsets, a tiny frame included."""
    cfg = CFGS[1]
    m, o = _make(cfg)
    rng = np.random.default_rng(5)
    frames = []
    for i, n in enumerate([2500, 40, 1200, 3100]):
        xyz = synth.surface_cloud(10 + i, bits=9, n_target=n) + rng.integers(0, 50, 3).astype(np.int32)
        frames.append(xyz[rng.permutation(xyz.shape[0])])
    want = [o.compress(synth.with_batch(f)) for f in frames]
    got = m.compress_batch([torch.from_numpy(synth.with_batch(f)).cuda() for f in frames])
    assert got == want
    rec = m.decompress_batch(got)
    for r, f, w in zip(rec, frames, want):
        assert (r.cpu().numpy() == o.decompress(w)).all()
        assert (np.unique(r.cpu().numpy(), axis=0) == np.unique(f, axis=0)).all()
    # partition container (3-byte length prefixes, model.py:455-463 / 510-521)
    blob = m.compress_partitions([None] + [torch.from_numpy(synth.with_batch(f)).cuda() for f in frames])
    assert blob == b''.join(len(s).to_bytes(3, 'little') + s for s in want)
    assert m.decompress_partitions(blob).shape[0] == sum(f.shape[0] for f in frames)


def test_concurrent_groups_on_a_fresh_model():
    """Groups run in threads on their own CUDA streams.  Derived parameter tensors (padded logits weights, occupancy
    bias tables, im2col weights) are built lazily on first use: a FRESH model coded with several groups at once must
    not let one stream read such a tensor before the stream that builds it has finished (they are published only
    after a stream synchronisation).  Bytes must equal the single-group result, repeatedly."""
    cfg = dict(channels=64, max_stride_wo_recurrent=32, max_stride=128, fea_stride=16)
    rng = np.random.default_rng(11)
    frames = []
    for i in range(9):
        xyz = synth.surface_cloud(30 + i, bits=9, n_target=3000 + 150 * i) + rng.integers(0, 50, 3).astype(np.int32)
        frames.append(torch.from_numpy(synth.with_batch(xyz)).cuda())
    ref_model, _ = _make(cfg)
    want = ref_model.compress_batch(frames)
    for trial in range(3):
        m, _ = _make(cfg)                      # fresh caches every trial
        got = m.compress_batch(frames, n_groups=3)
        assert got == want, trial
        rec = m.decompress_batch(got, n_groups=3)
        rec1 = ref_model.decompress_batch(want)
        assert all(torch.equal(a, b) for a, b in zip(rec, rec1))


def _drop_caches(model):
    """forget every lazily derived tensor (module caches and the process-wide ones) as a fresh process would"""
    from fastpcc_b200 import ops
    for mod in model.modules():
        for attr in ('_patch_weight', '_pad_cache', '_bits_cache', '_bit_levels', '_shift_cache', '_unit_range'):
            if hasattr(mod, attr):
                delattr(mod, attr)
    ops._POPC8.clear()
    ops._identity.clear()


@pytest.mark.parametrize('omp1', [False, True])
def test_concurrent_groups_first_use_under_host_contention(omp1):
    """Round-1 multi-GPU failure mode: under torchrun (OMP_NUM_THREADS=1, N processes on the same cores) a rank died in
    its first steps.  The first step of every coding group builds the shared derived tensors; with the host threads
    delayed at random that must still give the single-group bytes.  8 busy-loop threads hog the GIL and the cores
    while fresh caches are raced by 3 groups, 12 times."""
    import threading
    cfg = dict(channels=64, max_stride_wo_recurrent=32, max_stride=128, fea_stride=16)
    rng = np.random.default_rng(12)
    frames = []
    for i in range(9):
        xyz = synth.surface_cloud(50 + i, bits=9, n_target=2500 + 200 * i) + rng.integers(0, 50, 3).astype(np.int32)
        frames.append(torch.from_numpy(synth.with_batch(xyz)).cuda())
    m, _ = _make(cfg)
    want = m.compress_batch(frames)
    rec1 = m.decompress_batch(want)
    old_threads = torch.get_num_threads()
    if omp1:
        torch.set_num_threads(1)
    stop = threading.Event()

    def hog():
        x = 0
        while not stop.is_set():
            x = (x * 1103515245 + 12345) & 0x7fffffff

    hogs = [threading.Thread(target=hog, daemon=True) for _ in range(8)]
    for h in hogs:
        h.start()
    try:
        for trial in range(12):
            _drop_caches(m)
            got = m.compress_batch(frames, n_groups=3)
            assert got == want, trial
            _drop_caches(m)
            rec = m.decompress_batch(got, n_groups=3)
            assert all(torch.equal(a, b) for a, b in zip(rec, rec1)), trial
    finally:
        stop.set()
        for h in hogs:
            h.join()
        torch.set_num_threads(old_threads)


def test_bottom_level_too_wide_raises_on_the_host():
    """A bottom level wider than the 130-entry side-info table is a clean host error (reference: AssertionError at
    model.py:371), not a device-side index assert."""
    cfg = dict(channels=16, max_stride_wo_recurrent=16, max_stride=64, fea_stride=4, skip_top_scales_num=3)
    from fastpcc_b200.lossl_coord_int import Config, Model
    sd = synth.make_lossl_int_state_dict(seed=7, **{k: v for k, v in cfg.items() if k != 'skip_top_scales_num'})
    m = Model(Config(**cfg), device='cuda').load_numpy_state_dict(sd).cuda()
    xyz = synth.surface_cloud(3, bits=11, n_target=3000)  # 11-bit grid, 3 levels coded: bottom coordinates up to 255
    with pytest.raises(ValueError):
        m.compress(torch.from_numpy(synth.with_batch(xyz)).cuda())
    ok = synth.surface_cloud(3, bits=9, n_target=3000)
    assert len(m.compress(torch.from_numpy(synth.with_batch(ok)).cuda())) > 8  # the context is still healthy


def _golden():
    import json
    import os.path as osp
    from tests.golden.int_codec_cases import CASES
    gold = json.load(open(osp.join(osp.dirname(__file__), 'golden', 'int_codec_golden.json')))['cases']
    return list(zip(CASES, gold))


@pytest.mark.parametrize('case,gold', _golden(), ids=[c['name'] for c, _ in _golden()])
def test_bitstream_equals_reference_python_golden(case, gold):
    """The CUDA codec against bitstreams minted by the reference's own Python (unmodified model.py + cuda_ops.py +
    compiled range coder, run on the CPU by tests/golden/make_int_codec_golden.py): identical bytes, identical
    decoded points in the reference's output order.  Covers every block type and a 16-bit LiDAR pyramid."""
    import hashlib
    from fastpcc_b200.lossl_coord_int import Config, Model
    from tests.golden.int_codec_cases import case_cloud
    cfg = case['cfg']
    sd = synth.make_lossl_int_state_dict(seed=7, **{k: v for k, v in cfg.items() if k != 'skip_top_scales_num'})
    m = Model(Config(**cfg), device='cuda').load_numpy_state_dict(sd).cuda()
    xyz = case_cloud(case)
    data = m.compress(torch.from_numpy(synth.with_batch(xyz)).cuda())
    assert len(data) == gold['n_bytes']
    assert hashlib.sha256(data).hexdigest() == gold['bitstream_sha256']
    rec = m.decompress(data).cpu().numpy()
    assert hashlib.sha256(np.ascontiguousarray(rec.astype('<i4')).tobytes()).hexdigest() == gold['decoded_sha256']
