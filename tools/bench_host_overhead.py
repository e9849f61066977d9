"""Host-side overhead of one codec step: tiny frames (GPU work negligible) through the full Python path."""
import sys, time
import os.path as osp
import numpy as np
import torch
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from fastpcc_b200 import synth  # noqa: E402
from fastpcc_b200.lossl_coord_int import Config, Model  # noqa: E402

cfg = dict(channels=256, max_stride_wo_recurrent=2048, max_stride=8192, fea_stride=16)
sd = synth.make_lossl_int_state_dict(seed=7, **cfg)
m = Model(Config(**cfg), device='cuda').load_numpy_state_dict(sd).cuda()
frames = [torch.from_numpy(synth.with_batch(synth.lidar_frame(1000 + i)[::100])).cuda() for i in range(4)]
for _ in range(3):
    d = m.compress_batch(frames); r = m.decompress_batch(d)
torch.cuda.synchronize()
t = time.perf_counter(); d = m.compress_batch(frames); torch.cuda.synchronize(); t1 = time.perf_counter()
r = m.decompress_batch(d); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'tiny frames: compress {1e3 * (t1 - t):.1f} ms, decompress {1e3 * (t2 - t1):.1f} ms (host-bound)')
