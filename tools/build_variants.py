"""Builds the experiment variants of the library next to the default one and prints the single gpurun
command that checks parity of each and times it.  Run HERE (nvcc cross-compiles); the .so files travel with the repo
snapshot.  usage: python tools/build_variants.py"""
import os.path as osp
import sys

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastpcc_b200 import build  # noqa: E402

VARIANTS = {
    'pw2': ['-DFPCC_PAIRS_PRODUCER_WARPS=2'],
    'pw2_tile': ['-DFPCC_PAIRS_PRODUCER_WARPS=2', '-DFPCC_EPI_TILE_DISPATCH=1'],
    'pw2_pair': ['-DFPCC_PAIRS_PRODUCER_WARPS=2', '-DFPCC_EPI_PAIR_STORES=1'],
}

if __name__ == '__main__':
    build.build()
    cmds = ['python tools/profile_calls.py 32 > gpurun_out/calls_base.txt 2>&1']
    for name, flags in VARIANTS.items():
        rel = osp.relpath(build.build_variant(name, flags), ROOT)
        env = f'FPCC_LIB_PATH=$PWD/{rel}'
        cmds.append(f'{env} timeout 120 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q > gpurun_out/pytest_{name}.log 2>&1; '
                    f'{env} python tools/profile_calls.py 32 > gpurun_out/calls_{name}.txt 2>&1')
    cmds.append('grep -H "==" gpurun_out/calls_*.txt; tail -n 2 gpurun_out/pytest_*.log')
    print("gpurun --timeout 600 -- '" + '; '.join(cmds) + "'")
