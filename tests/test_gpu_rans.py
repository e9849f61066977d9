"""GPU range coders: byte-identical to the golden vectors minted from the compiled reference, and to the
CPU oracle on random streams (several streams per launch, ragged lengths, empty stream)."""
import numpy as np
import pytest
import torch

from oracle import rans as orans
from tests.golden import rans_cases

pytestmark = pytest.mark.gpu


def test_gpu_coders_match_reference_golden(rans_kat):
    from fastpcc_b200 import rans_coder as g
    got = rans_cases.run_all(g.RansEncoder, g.RansDecoder, g.IndexedRansCoder, g.BinaryRansCoder,
                             g.batched_pmf_to_quantized_cdf)
    for k, v in rans_kat.items():
        if not k.startswith('_'):
            assert got[k] == v, k


def test_batched_streams_vs_oracle():
    from fastpcc_b200 import ops
    rng = np.random.default_rng(0)
    lens = [0, 1, 31, 32, 33, 5000, 12345]
    S = 255
    cdfs, syms, want = [], [], []
    for n in lens:
        pm = rng.integers(1, 60, (n, S)) ** 2
        pm = pm * (65536 - S) // np.maximum(pm.sum(1, keepdims=True), 1) + 1
        cdf = np.cumsum(pm, 1); cdf[:, -1] = 65535
        cdf = cdf.astype(np.uint16)
        sym = rng.integers(0, S, n).astype(np.uint16)
        e = orans.RansEncoder(1 << 20)
        if n:
            e.encode(cdf, sym)
        want.append(e.flush())
        cdfs.append(cdf); syms.append(sym)
    cdf_all = np.concatenate(cdfs)
    sym_all = np.concatenate(syms).astype(np.int32)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    d_cdf = torch.from_numpy(cdf_all).cuda()
    ranges = ops.table_symbol_ranges(d_cdf, torch.from_numpy(sym_all).cuda())
    cap = 2 * max(lens) + 64
    out, out_len = ops.rans_encode(ranges, torch.from_numpy(off).cuda(), cap)
    out, out_len = out.cpu().numpy(), out_len.cpu().numpy()
    got = [out[b, cap - out_len[b]:].tobytes() for b in range(len(lens))]
    assert got == want
    # decode all streams in one launch from a padded (ld=256) table: the fast path
    pad = np.full((cdf_all.shape[0], 256), 0xFFFF, np.uint16); pad[:, :S] = cdf_all
    blob = np.frombuffer(b''.join(got), np.uint8).copy()
    boff = np.concatenate([[0], np.cumsum([len(g) for g in got])]).astype(np.int64)
    dec = ops.RansDecodeStreams(torch.from_numpy(blob).cuda(), torch.from_numpy(boff[:-1].copy()).cuda(),
                                torch.tensor([len(g) for g in got], dtype=torch.int32).cuda())
    sym = dec.decode(torch.from_numpy(pad).cuda(), S, torch.from_numpy(off).cuda(), int(off[-1]))
    assert (sym.cpu().numpy() == sym_all).all() and not dec.error()
    # and through the generic path (unpadded rows)
    dec2 = ops.RansDecodeStreams(torch.from_numpy(blob).cuda(), torch.from_numpy(boff[:-1].copy()).cuda(),
                                 torch.tensor([len(g) for g in got], dtype=torch.int32).cuda())
    sym2 = dec2.decode(d_cdf, S, torch.from_numpy(off).cuda(), int(off[-1]))
    assert (sym2.cpu().numpy() == sym_all).all()


def test_truncated_stream_is_reported():
    from fastpcc_b200 import rans_coder as g
    cdf, sym = rans_cases.kat_e_inputs(2000)
    enc = g.RansEncoder(1 << 20)
    enc.encode(cdf, sym)
    data = enc.flush()
    dec = g.RansDecoder()
    dec.flush(data[: len(data) // 2])
    out = np.zeros_like(sym)
    with pytest.raises(RuntimeError):
        dec.decode(cdf, out)


def test_lossy_heads_match_oracle_and_roundtrip():
    """geo_lossl_em.py:59-114: sigmoid -> 16-bit P(1) -> binary rANS; histogram-CDF residual stream layout."""
    import io
    from fastpcc_b200 import lossy_heads as H
    g = torch.Generator(device='cuda').manual_seed(0)
    logit = (torch.randn((50000, 1), generator=g, device='cuda') * 4).half()
    x = (torch.rand(50000, generator=g, device='cuda') < logit.float().sigmoid().reshape(-1))
    prob = np.clip(np.round(logit.sigmoid().cpu().numpy().astype(np.float64) * 65536).astype(np.uint32), 1, 65535).reshape(1, -1)
    assert (H.init_prob(logit).cpu().numpy().astype(np.uint32) == prob.reshape(-1)).all()
    want = orans.BinaryRansCoder(1).encode(x.cpu().numpy().reshape(1, -1), prob)[0]
    data = H.binary_encode(logit, x)
    assert data == want
    assert torch.equal(H.binary_decode(logit, data), x)
    rng = np.random.default_rng(1)
    t = np.round(np.clip(rng.normal(0, 3, (4000, 2)), -20, 20)).astype(np.int32)
    bs = io.BytesIO()
    H.rans_encode_with_cdf(t, bs)
    ref = io.BytesIO()
    # the same layout through the oracle coder
    ref.write((4000).to_bytes(3, 'little')); off = int(t.min()); ref.write((-off).to_bytes(1, 'little'))
    oc = orans.IndexedRansCoder(False, 1)
    oc.init_with_pmfs(np.bincount((t - off).reshape(-1)).astype(np.float64)[None], np.array([off], np.int32))
    cdf = oc.get_cdfs()[0]
    ref.write((len(cdf) - 2).to_bytes(1, 'little'))
    for cd in cdf[1:-1]:
        ref.write(int(cd).to_bytes(2, 'little'))
    payload = oc.encode(t.reshape(1, -1))[0]
    ref.write(len(payload).to_bytes(3, 'little')); ref.write(payload)
    assert bs.getvalue() == ref.getvalue()
    bs.seek(0)
    back, cdf2 = H.rans_decode_with_cdf(bs, channels=2)
    assert (back == t).all() and cdf2 == cdf


def test_lossy_topk_keep_mask():
    """lossy_coord_v2/layers.py:151-180 restated with numpy loops: group max always kept, k-th value threshold."""
    from fastpcc_b200 import lossy_heads as H
    rng = np.random.default_rng(5)
    coords, tgt = [], []
    for b in range(2):
        xyz = np.unique(rng.integers(0, 13, (1500, 3)), axis=0) * 2          # stride-2 candidates, ~40 per stride-8 group
        coords.append(np.concatenate([np.full((len(xyz), 1), b), xyz], 1))
        tgt.append(len(xyz) // 3)
    C = np.concatenate(coords).astype(np.int32)
    f = rng.normal(size=len(C)).astype(np.float32)
    keep = H.get_keep(torch.from_numpy(f).cuda(), torch.from_numpy(C).cuda(), [2, 2, 2], [8, 8, 8], list(tgt)).cpu().numpy()
    groups = {}
    for i, (b, x, y, z) in enumerate(C):
        groups.setdefault((b, x // 8, y // 8, z // 8), []).append(i)
    is_max = np.zeros(len(C), bool)
    for idx in groups.values():
        m = f[idx].max()
        for i in idx:
            is_max[i] = f[i] == m
    want = is_max.copy()
    for b in range(2):
        rows = C[:, 0] == b
        cand = np.sort(f[rows & ~is_max])
        thr = cand[rows.sum() - tgt[b] - 1]
        want |= rows & (f > thr)
    assert (keep == want).all()
    keep0 = H.get_keep(torch.from_numpy(f).cuda(), torch.from_numpy(C).cuda(), [2, 2, 2], [8, 8, 8]).cpu().numpy()
    assert (keep0 == (is_max | (f > 0))).all()
