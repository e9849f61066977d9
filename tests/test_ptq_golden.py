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
