"""CPU oracle for the sparse-convolution codec hot path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this package; `fastpcc_b200` never does.

  rans.py            ctypes front-end of rans_oracle.c (plain-C restatement of the reference coders)
  int_ops.py         numpy restatement of lib/int_sparse_conv (kernel map, int8 conv/linear, requant, softmax)
  lossl_coord_int.py numpy restatement of models/convolutional/lossl_coord_int/model.py
  build_ref.py       compiles the reference's own coders from /root/reference into oracle/_ref
"""
import os.path as osp
import subprocess

HERE = osp.dirname(osp.abspath(__file__))


def build(verbose=False):
    """Compile the C restatement (gcc) and, when /root/reference is present, oracle/_ref."""
    subprocess.run(['make', '-C', HERE], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    from . import build_ref
    try:
        build_ref.build(verbose=verbose)
    except Exception as e:  # the reference build is a strengthening, not a requirement
        print(f'[oracle] reference build skipped: {e}')
