"""Per-call device times of one compress + decompress batch (diagnostics).  usage: python tools/profile_calls.py [frames=8]"""
import sys
import os.path as osp
import torch
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from fastpcc_b200 import ops, synth  # noqa: E402
from fastpcc_b200.lossl_coord_int import Config, Model  # noqa: E402

cfg = dict(channels=256, max_stride_wo_recurrent=2048, max_stride=8192, fea_stride=16)
m = Model(Config(**cfg), device='cuda').load_numpy_state_dict(synth.make_lossl_int_state_dict(seed=7, **cfg)).cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
frames = [torch.from_numpy(synth.with_batch(synth.lidar_frame(1000 + i))).cuda() for i in range(B)]
m.decompress_batch(m.compress_batch(frames))
for phase in ('compress', 'decompress'):
    prof = ops.enable_profile(True)
    if phase == 'compress':
        d = m.compress_batch(frames)
    else:
        m.decompress_batch(d)
    torch.cuda.synchronize()
    rows = []
    for tag, e0, e1, w in prof:
        ms = e0.elapsed_time(e1)
        rate = f"{w['bytes'] / ms / 1e6:7.0f} GB/s " if w.get('bytes') else (f"{w['ops'] / ms / 1e9:7.0f} TOP/s" if w.get('ops') else ' ' * 13)
        rows.append((ms, tag, rate + ' ' + str(w.get('desc', ''))))
    ops.enable_profile(False)
    print(f'== {phase}: {len(rows)} calls, {sum(r[0] for r in rows):.1f} ms in kernels')
    for ms, tag, desc in sorted(rows, reverse=True)[:80]:
        print(f'  {ms:8.3f} ms  {tag:28s} {desc}')
