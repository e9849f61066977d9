"""PTQ end to end on the device: a float subtree (residual block + conv/PReLU/linear head) runs on the fp16 tcgen05
kernels with observers in place, is converted by fastpcc_b200/ptq.py, and the integer modules (int8 tcgen05 kernels,
Q8.23 activations) reproduce the float output within the quantisation error of 8-bit activations and weights."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from fastpcc_b200 import synth

pytestmark = pytest.mark.gpu


def test_calibrate_convert_and_compare_with_float():
    from fastpcc_b200 import ptq, torchsparse_nn as TS
    from fastpcc_b200.int_sparse_conv.cuda_ops import SharedFxpShift
    from fastpcc_b200.sparse_tensor import SparseTensor
    torch.manual_seed(3)
    ch = 32
    C = synth.with_batch(synth.surface_cloud(6, bits=7, n_target=5000))
    C = torch.from_numpy(C[np.lexsort((C[:, 3], C[:, 2], C[:, 1], C[:, 0]))]).cuda()
    f = torch.randn(C.shape[0], ch, device='cuda') * 0.7

    model = nn.ModuleDict({
        'dec': TS.Block(ch),
        'pred': TS.SparseSequential(TS.Conv3d(ch, ch, 3, 1, 1, bias=True), nn.PReLU(), nn.Linear(ch, 48)),
    }).cuda()

    def run(m, feats):
        x = SparseTensor(feats, C, 1)
        return m['pred'](m['dec'](x)).F

    with torch.no_grad():
        ptq.insert_observers(model)
        assert isinstance(model['dec'], ptq.SparseResBlockWithObs) and len(model['pred']) == 6
        want = run(model, f.half()).float()          # the calibration pass: observers see every activation
        ptq.convert_to_int(model)
        names = [type(m).__name__ for m in model['pred']]
        assert names == ['RequantFxpToScaledInt8', 'SparseConvPReLUIn8W8Out8', 'LinearIn8W8Out32'], names
        assert type(model['dec']).__name__ == 'SparseResBlockIn32W8Out32'
        fxp = (f * (1 << SharedFxpShift)).round().to(torch.int32)
        got = run(model, fxp)
        assert got.dtype == torch.int32
        got = got.double() / (1 << SharedFxpShift)
    err = (got - want.double()).norm() / want.double().norm()
    assert err < 0.05, float(err)   # int8 activations + int8 per-channel weights over three quantised layers
    assert err > 0                  # and it is the integer path that ran, not the float one
