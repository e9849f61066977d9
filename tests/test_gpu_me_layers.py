"""Float layer API (twin of lib/minkowski_sparse_conv_layers.py) on the tcgen05 fp16 kernels vs an fp32 numpy
restatement of the MinkowskiEngine semantics (oracle/float_ops.py).  Tolerance: 1e-2 relative per layer in fp16
(BASELINE.json), 3e-2 for a 5-layer stack."""
import numpy as np
import pytest
import torch

from fastpcc_b200 import synth
from oracle import float_ops as FO

pytestmark = pytest.mark.gpu


def _rel(got, want):
    return float(np.abs(got.astype(np.float64) - want).max() / max(np.abs(want).max(), 1e-12))


def _np(t):
    return t.detach().float().cpu().numpy()


def _sorted_cloud(ME, seed, n, bits):
    rng = np.random.default_rng(seed)
    pts = np.unique(rng.integers(0, 1 << bits, (n, 3)), axis=0).astype(np.int32)
    C = torch.from_numpy(synth.with_batch(pts)).cuda()
    return C[ME.morton_order(C)].contiguous()


def _conv_ref(x_f, C_in, C_out_coords, blk, ks, scale, act=None, slope=0.0):
    kern = _np(blk.conv.kernel)
    kern = kern[None] if kern.ndim == 2 else kern
    table = FO.me_lookup(C_in, C_out_coords, ks, scale)
    bias = _np(blk.conv.bias).reshape(-1) if blk.conv.bias is not None else None
    return FO.act(FO.sparse_conv_f32(x_f, kern, table, bias), act, slope)


def test_blocks_match_fp32_reference():
    from fastpcc_b200 import me as ME
    from fastpcc_b200 import minkowski_sparse_conv_layers as L
    torch.manual_seed(0)
    C = _sorted_cloud(ME, 0, 6000, 6)
    n = C.shape[0]
    x = ME.SparseTensor(torch.ones((n, 1), device='cuda'), coordinates=C, tensor_stride=1)
    Cn = C.cpu().numpy()

    b0 = L.ConvBlock(1, 16, 3, 1, act='relu').cuda()
    y0 = b0(x)
    r0 = _conv_ref(np.ones((n, 1), np.float32), Cn, Cn, b0, 3, 1, 'relu')
    assert y0.F.dtype == torch.float16 and _rel(_np(y0.F), r0) < 1e-2

    b1 = L.ConvBlock(16, 64, 2, 2, act='leaky_relu(0.2)').cuda()
    y1 = b1(y0)
    oc = Cn.copy(); oc[:, 1:] = oc[:, 1:] // 2 * 2
    oc = np.unique(oc, axis=0)
    C1 = y1.C.cpu().numpy()
    assert y1.tensor_stride == [2, 2, 2] and C1.shape == oc.shape
    assert (np.unique(C1, axis=0) == oc).all()
    r1 = _conv_ref(_np(y0.F), Cn, C1, b1, 2, 1, 'leaky_relu', 0.2)
    assert _rel(_np(y1.F), r1) < 1e-2

    rb = L.ResBlock(64, 'HYPER_CUBE', False, 'relu', last_act=True).cuda()
    y2 = rb(y1)
    t = _conv_ref(_np(y1.F), C1, C1, rb.conv0, 3, 2, 'relu')
    t = _conv_ref(t, C1, C1, rb.conv1, 3, 2) + _np(y1.F)
    assert _rel(_np(y2.F), np.maximum(t, 0)) < 2e-2

    ib = L.InceptionResBlock(64, 'HYPER_CUBE', False, 'relu').cuda()
    y3 = ib(y2)
    f2 = _np(y2.F)
    p0 = _conv_ref(_conv_ref(f2, C1, C1, ib.path_0[0], 3, 2, 'relu'), C1, C1, ib.path_0[1], 3, 2)
    p1 = _conv_ref(f2, C1, C1, ib.path_1[0], 1, 2, 'relu')
    p1 = _conv_ref(_conv_ref(p1, C1, C1, ib.path_1[1], 3, 2, 'relu'), C1, C1, ib.path_1[2], 1, 2)
    assert _rel(_np(y3.F), np.concatenate([p0, p1], 1) + f2) < 2e-2

    # generative transposed conv: all 8 children of every stride-2 voxel, rows (parent, child) with x fastest
    gb = L.GenConvTransBlock(64, 16, 2, 2, act='relu').cuda()
    y4 = gb(y3)
    C4 = y4.C.cpu().numpy()
    exp = (L.minkowski_expand_coord_2x(torch.from_numpy(C1), 2)).reshape(-1, 4).numpy()
    assert y4.tensor_stride == [1, 1, 1] and (C4 == exp).all()
    kern = _np(gb.conv.kernel)
    f3 = _np(y3.F)
    want = np.stack([f3 @ kern[k] for k in range(8)], 1).reshape(-1, 16) + _np(gb.conv.bias)
    assert _rel(_np(y4.F), np.maximum(want, 0)) < 1e-2

    # pruning back to the true voxels, then a transposed conv of the stride-2 features onto that key
    lut = {tuple(c): i for i, c in enumerate(Cn.tolist())}
    mask = torch.tensor([tuple(c) in lut for c in C4.tolist()], device='cuda')
    y5 = ME.MinkowskiPruning()(y4, mask)
    assert y5.F.shape[0] == n
    tb = L.ConvTransBlock(64, 32, 2, 2, act=None).cuda()
    y6 = tb(y3, y5.coordinate_map_key)
    C5 = y5.C.cpu().numpy()
    par = {tuple(c): i for i, c in enumerate(C1.tolist())}
    kt = _np(tb.conv.kernel)
    want = np.zeros((n, 32), np.float32)
    for i, c in enumerate(C5.tolist()):
        p = (c[0], c[1] // 2 * 2, c[2] // 2 * 2, c[3] // 2 * 2)
        k = (c[1] & 1) + 2 * (c[2] & 1) + 4 * (c[3] & 1)
        want[i] = f3[par[p]] @ kt[k]
    want += _np(tb.conv.bias)
    assert _rel(_np(y6.F), want) < 1e-2

    mlp = L.MEMLPBlock(32, 1, act=None).cuda()
    y7 = mlp(y6)
    want7 = _np(y6.F) @ _np(mlp.mlp.linear.weight).T + _np(mlp.mlp.linear.bias)
    assert y7.F.shape == (n, 1) and _rel(_np(y7.F), want7) < 1e-2
    # kernel_map(kernel_size=1) between a pruned key and its superset (get_coord_mask, geo_lossl_em.py:306-317)
    km = y4.coordinate_manager.kernel_map(y5.coordinate_map_key, y4.coordinate_map_key, kernel_size=1)
    assert list(km) == [0] and km[0].shape == (2, n)
    assert (km[0][1].cpu() == torch.nonzero(mask.cpu())[:, 0]).all()


def test_max_pool_unpool_and_per_sample_views_drive_topk_pruning():
    """The lossy decoder's top-k pruning (lossy_coord_v2/layers.py:151-180) written against the ME surface --
    MinkowskiMaxPooling / MinkowskiPoolingTranspose onto the coarse key, decomposition_permutations, kthvalue --
    must select the same voxels as lossy_heads.get_keep (scatter-reduce formulation) and as a numpy loop."""
    import numpy as np
    from fastpcc_b200 import me as ME, lossy_heads as H
    rng = np.random.default_rng(3)
    cs = []
    for b in range(2):
        xyz = np.unique(rng.integers(0, 13, (1500, 3)), axis=0) * 2
        cs.append(np.concatenate([np.full((len(xyz), 1), b), xyz], 1))
    C = torch.from_numpy(np.concatenate(cs).astype(np.int32)).cuda()
    C = C[ME.morton_order(C, 2)].contiguous()
    f = torch.from_numpy(rng.normal(size=(C.shape[0], 1)).astype(np.float32)).cuda()
    pred = ME.SparseTensor(f, C, tensor_stride=2)
    cm = pred.coordinate_manager
    coarse = cm.stride(pred.coordinate_map_key, 4)                       # stride 8 cells
    assert len(cm._manager.get_coordinate_map_keys([8, 8, 8])) == 1
    pool, unpool = ME.MinkowskiMaxPooling(4, 4, dimension=3), ME.MinkowskiPoolingTranspose(4, 4, dimension=3)
    local_max = unpool(pool(pred, coarse), pred.coordinate_map_key)
    not_max = (pred.F - local_max.F).squeeze(1) != 0
    targets = [int(r.shape[0]) // 3 for r in pred.decomposition_permutations]
    assert [c.shape[0] for c in pred.decomposed_coordinates] == [int((C[:, 0] == b).sum()) for b in range(2)]
    thr = []
    for tgt, rows in zip(targets, pred.decomposition_permutations):
        sample = pred.F[rows]
        thr.append(torch.kthvalue(sample[not_max[rows]], sample.shape[0] - tgt, dim=0).values)
    thr = torch.cat(thr)[pred.C[:, 0].long()]
    keep = (pred.F.squeeze(1) > thr) | ~not_max
    assert torch.equal(keep, H.get_keep(pred.F, pred.C, [2, 2, 2], [8, 8, 8], targets))
    # numpy restatement of the cell maximum
    Cn, fn = C.cpu().numpy(), f.cpu().numpy()[:, 0]
    cell = {}
    for i, (b, x, y, z) in enumerate(Cn):
        cell.setdefault((b, x // 8, y // 8, z // 8), []).append(i)
    want = np.zeros(len(Cn), bool)
    for idx in cell.values():
        want[idx] = fn[idx] != fn[idx].max()
    assert (not_max.cpu().numpy() == want).all()
    pruned = ME.MinkowskiPruning()(pred, keep)
    assert pruned.coordinate_map_key.tag == 'pruned' and pruned.F.shape[0] == int(keep.sum())
