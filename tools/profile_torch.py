"""Whole-GPU kernel time table (ours + torch's own kernels) of one compress + decompress batch via torch.profiler, with the
wall time of each phase next to the summed device time (the difference is host-side gaps).
usage: python tools/profile_torch.py [frames=32]"""
import sys
import time
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
# un-profiled wall times first
t0 = time.perf_counter(); d = m.compress_batch(frames); torch.cuda.synchronize(); t1 = time.perf_counter()
m.decompress_batch(d); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'wall: compress {1e3 * (t1 - t0):.1f} ms, decompress {1e3 * (t2 - t1):.1f} ms ({B} frames, one group)')
for phase in ('compress', 'decompress'):
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        if phase == 'compress':
            d = m.compress_batch(frames)
        else:
            m.decompress_batch(d)
        torch.cuda.synchronize()
    ka = prof.key_averages()
    kern = [e for e in ka if e.device_type == torch.autograd.DeviceType.CUDA or (e.self_device_time_total > 0 and e.self_cpu_time_total == 0)]
    tot = sum(e.self_device_time_total for e in kern)
    ours = sum(e.self_device_time_total for e in kern if 'fpcc' in e.key)
    print(f'== {phase}: device time {tot / 1e3:.1f} ms in {sum(e.count for e in kern)} kernels/copies (fpcc kernels {ours / 1e3:.1f} ms)')
    for e in sorted(kern, key=lambda e: -e.self_device_time_total)[:28]:
        print(f'{e.self_device_time_total / 1e3:9.2f} ms  x{e.count:5d}  {e.key[:120]}')
    cpu = sorted(ka, key=lambda e: -e.self_cpu_time_total)[:12]
    print('   top host-side ops (self CPU time):')
    for e in cpu:
        print(f'{e.self_cpu_time_total / 1e3:9.2f} ms  x{e.count:5d}  {e.key[:100]}')
