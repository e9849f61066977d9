"""Whole-GPU kernel time table (ours + torch's own kernels) of one compress + decompress batch via torch.profiler.
usage: python tools/profile_torch.py [frames=32]"""
import sys
import os.path as osp
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from fastpcc_b200 import synth  # noqa: E402
from fastpcc_b200.lossl_coord_int import Config, Model  # noqa: E402

cfg = dict(channels=256, max_stride_wo_recurrent=2048, max_stride=8192, fea_stride=16)
m = Model(Config(**cfg), device='cuda').load_numpy_state_dict(synth.make_lossl_int_state_dict(seed=7, **cfg)).cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
frames = [torch.from_numpy(synth.with_batch(synth.lidar_frame(1000 + i))).cuda() for i in range(B)]
m.decompress_batch(m.compress_batch(frames))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    d = m.compress_batch(frames)
    torch.cuda.synchronize()
    m.decompress_batch(d)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.self_device_time_total for e in ev)
print(f'total device time {tot / 1e3:.1f} ms')
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:40]:
    if e.self_device_time_total > 0:
        print(f'{e.self_device_time_total / 1e3:9.2f} ms  x{e.count:5d}  {e.key[:110]}')
