"""The PTQ import formulas of the layer API twin (fastpcc_b200/int_sparse_conv/cuda_ops.py `import_parameters`)
against the reference's own classes (lib/int_sparse_conv/cuda_ops.py:223-301, 464-468, 488-501, 542-600) run on the
same seeded float layers: tests/golden/ptq_golden.json is minted by tests/golden/make_int_codec_golden.py from the
unmodified reference Python.  Every integer buffer (weights, bias, slope, requant multiplier / shift, zero points,
zero-point compensation rows) must be identical, dtype and shape included.  Host logic only: runs without a GPU."""
import json
import os.path as osp

from tests.golden.ptq_cases import run_cases

GOLDEN = json.load(open(osp.join(osp.dirname(__file__), 'golden', 'ptq_golden.json')))


def test_import_parameters_matches_reference_classes():
    from fastpcc_b200.int_sparse_conv import cuda_ops
    got = run_cases(cuda_ops)
    assert sorted(got) == sorted(GOLDEN)
    for name, bufs in GOLDEN.items():
        assert got[name] == bufs, name


def test_convert_pipeline_matches_reference_functions():
    """insert_obs_into_* / replace_*_with_int_impl of fastpcc_b200/ptq.py on the float predictor trees against the
    reference's own functions on the reference's own trees (tests/golden/ptq_pipeline_golden.json): same state-dict
    keys, same integer buffers -- i.e. the observer placement (affine before Linear, symmetric otherwise), the
    PReLU fusion, the scaled-int8 / Q8.23 hand-over between layers and every requant parameter agree."""
    import types
    from fastpcc_b200 import ptq
    from fastpcc_b200.sparse_tensor import SparseTensor
    from tests.golden.ptq_cases import run_pipeline
    gold = json.load(open(osp.join(osp.dirname(__file__), 'golden', 'ptq_pipeline_golden.json')))
    ns = types.SimpleNamespace(
        OneScalePredictor=ptq.OneScalePredictor, OneScaleMultiStepPredictor=ptq.OneScaleMultiStepPredictor,
        insert_obs_into_resblocks=ptq.insert_obs_into_resblocks, insert_obs_into_seqs=ptq.insert_obs_into_seqs,
        replace_resblocks_with_int_impl=ptq.replace_resblocks_with_int_impl, replace_seqs_with_int_impl=ptq.replace_seqs_with_int_impl,
        SparseTensorHistogramObserver=ptq.SparseTensorHistogramObserver, SparseTensor=SparseTensor)
    got = run_pipeline(ns)
    assert sorted(got) == sorted(gold)
    bad = [k for k in gold if got[k] != gold[k]]
    assert not bad, bad[:10]


def test_converted_trunk_has_the_integer_models_state_dict_keys():
    """A converted float parameter tree must load into lossl_coord_int.Model (same keys, dtypes, shapes)."""
    import torch
    from fastpcc_b200 import ptq
    from fastpcc_b200.lossl_coord_int import Config, Model
    from fastpcc_b200.sparse_tensor import SparseTensor
    cfg = dict(channels=16, max_stride_wo_recurrent=32, max_stride=128, fea_stride=16)
    trunk = ptq.FloatTrunk(**cfg)
    ptq.insert_observers(trunk)
    for m in trunk.modules():
        if isinstance(m, ptq.SparseTensorHistogramObserver):
            m(SparseTensor(torch.randn(256, 4), torch.zeros((256, 4), dtype=torch.int32)))
    ptq.convert_to_int(trunk)
    sd = {k: v for k, v in trunk.state_dict().items()}
    want = {k: v for k, v in Model(Config(**cfg), device='cpu').state_dict().items()}
    assert sorted(sd) == sorted(want)
    for k in want:
        assert sd[k].dtype == want[k].dtype and sd[k].shape == want[k].shape, k
