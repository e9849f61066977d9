"""Mints tests/golden/container_golden.json with the reference's own BytesListUtils
(lib/entropy_models/hyperprior/noisy_deep_factorized/utils.py, pure Python, loaded from its file unmodified).
Run:  python tests/golden/make_container_golden.py"""
import importlib.util
import json
import os.path as osp
import sys

HERE = osp.dirname(osp.abspath(__file__))
sys.path.insert(0, osp.dirname(osp.dirname(HERE)))
from tests.golden.container_cases import bytes_lists  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location(
        'ref_bl_utils', '/root/reference/lib/entropy_models/hyperprior/noisy_deep_factorized/utils.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    for name, items in bytes_lists().items():
        blob = mod.BytesListUtils.concat_bytes_list(items)
        assert mod.BytesListUtils.split_bytes_list(blob, len(items)) == items
        import hashlib
        out[name] = {'n': len(items), 'len': len(blob), 'head_hex': blob[:64].hex(), 'sha256': hashlib.sha256(blob).hexdigest()}
        print(name, len(items), len(blob))
    with open(osp.join(HERE, 'container_golden.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
