"""Mints tests/golden/lossy_v2_golden.json: the reference's OWN, unmodified lossy_coord_v2 codec
(models/convolutional/lossy_coord_v2/model.py + layers.py, lossy_coord_lossy_color/geo_lossl_em.py,
lib/minkowski_sparse_conv_layers.py, the reference's compiled range coders) run in this container on the CPU
MinkowskiEngine stand-in of the oracle (oracle/me_cpu.py: fp32 operators restated in oracle/float_ops_cpu.py).

MinkowskiEngine is not installable offline, so this is NOT the reference's arithmetic backend: the fixture pins what
the reference's model code produces on an fp32 restatement of ME's operators -- rate (bytes, bpp), D1 PSNR, and the
losslessly coded stride-2 geometry -- for the fp16 tensor-core path to be compared with (tests/test_gpu_lossy_dropin.py).

Run:  python tests/golden/make_lossy_golden.py     (needs /root/reference and oracle/_ref)
"""
import hashlib
import json
import os.path as osp
import sys

import numpy as np
import torch

HERE = osp.dirname(osp.abspath(__file__))
ROOT = osp.dirname(osp.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref, me_cpu, metrics  # noqa: E402
from fastpcc_b200 import synth  # noqa: E402
from tests import ref_import  # noqa: E402
from tests.golden.lossy_cases import CASES, case_cloud  # noqa: E402


def stride2_sha(xyz, origin):
    """occupied stride-2 cells on the codec's grid (coordinates relative to the cloud's minimum, model.py:231-232)"""
    c = np.unique((xyz - origin) // 2, axis=0)
    return hashlib.sha256(np.ascontiguousarray(c.astype('<i4')).tobytes()).hexdigest(), int(c.shape[0])


def run_case(ref, case, device='cpu'):
    cfg = dict(ref_import.LOSSY_V2_BASELINE_R1, bottleneck_scaler=case['bottleneck_scaler'])
    model = ref_import.build_reference_lossy_v2(ref, cfg, seed=case['model_seed'], device=device)
    xyz = case_cloud(case)
    c = torch.from_numpy(synth.with_batch(xyz)).to(device)
    with torch.no_grad():
        data = model.compress(c)
        rec = model.decompress(data).cpu().numpy()
    return xyz, data, rec


def main():
    me = me_cpu.load('MinkowskiEngine_cpu_oracle')
    ref = ref_import.import_reference_lossy_v2('/root/reference', me, rans=build_ref.load_ref('rans_ext_cpp'))
    out = {'_doc': 'reference lossy_coord_v2 model code on the fp32 CPU ME stand-in (see make_lossy_golden.py)', 'cases': []}
    for case in CASES:
        xyz, data, rec = run_case(ref, case)
        err = metrics.pc_error(xyz, rec, 2 ** case['bits'])
        sha, n2 = stride2_sha(xyz, xyz.min(0))
        out['cases'].append({'name': case['name'], 'n_points': int(xyz.shape[0]), 'n_bytes': len(data),
                             'bpp': len(data) * 8 / xyz.shape[0], 'n_rec': int(rec.shape[0]),
                             'd1_psnr': err['mseF,PSNR (p2point)'], 'd1_mse1': err['mse1      (p2point)'],
                             'd1_mse2': err['mse2      (p2point)'], 'stride2_sha256': sha, 'stride2_points': n2})
        print(out['cases'][-1])
    with open(osp.join(HERE, 'lossy_v2_golden.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
