"""Mints tests/golden/rans_kat.json by RUNNING THE REFERENCE's own coders (oracle/_ref, compiled
from /root/reference by oracle/build_ref.py).  Run in the build container only:
    python tests/golden/make_rans_kat.py
The cases are defined in tests/golden/rans_cases.py, shared with the tests that replay them.
"""
import hashlib
import json
import os.path as osp
import sys

import numpy as np

ROOT = osp.dirname(osp.dirname(osp.dirname(osp.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from tests.golden import rans_cases  # noqa: E402


def main():
    assert build_ref.build(), 'reference sources not available'
    simple = build_ref.load_ref('simple_rans_ext_cpp')
    batched = build_ref.load_ref('rans_ext_cpp')
    out = rans_cases.run_all(simple.RansEncoder, simple.RansDecoder, batched.IndexedRansCoder,
                             batched.BinaryRansCoder, batched.batched_pmf_to_quantized_cdf)
    out['_meta'] = {'minted_from': 'oracle/_ref (reference C++ compiled from /root/reference)',
                    'numpy': np.__version__}
    with open(osp.join(osp.dirname(osp.abspath(__file__)), 'rans_kat.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, str(v)[:100])


if __name__ == '__main__':
    main()
