"""DDP training step of the float sparse-conv path: one process per GPU, NCCL gradient all-reduce only.
usage: python tools/bench_train.py [--steps 10]         (1 GPU)
       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py
Prints one JSON line on rank 0: nodes per second over all ranks (weak scaling: every rank trains its own scans)."""
import argparse
import json
import os
import os.path as osp
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from fastpcc_b200 import autograd, ops, train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--warmup', type=int, default=3)
ap.add_argument('--frames', type=int, default=4)
ap.add_argument('--channels', type=int, default=128)
ap.add_argument('--blocks', type=int, default=2)
ap.add_argument('--wgrad', default='tc', choices=['tc', 'lib'])
args = ap.parse_args()
rank, local, world = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
autograd.WGRAD_TC = args.wgrad == 'tc'
torch.manual_seed(0)
model = train.OccupancyNet(args.channels, args.blocks).to(dev)
ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
opt = torch.optim.Adam(ddp.parameters(), lr=1e-3)
batch = train.make_batch([5000 + rank * args.frames + i for i in range(args.frames)], 4, dev)
n_nodes = batch[1].shape[0]
losses = []
for _ in range(args.warmup):
    losses.append(float(train.train_step(ddp, opt, batch)))
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
prof = ops.enable_profile(True) if rank == 0 else None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = train.train_step(ddp, opt, batch)
e1.record()
e1.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps, float(n_nodes)], dtype=torch.float64, device=dev)
if world > 1:
    mx = ms.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(ms, op=dist.ReduceOp.SUM)
    ms[0] = mx[0]
losses.append(float(loss))
if rank == 0:
    stats = ops.profile_summary(prof)
    ops.enable_profile(False)
    k = {t: {'ms_per_step': round(v['ms'] / args.steps, 3), 'tflops': round(v['ops'] / (v['ms'] * 1e-3) / 1e12, 1) if v.get('ops') and v['ms'] else None}
         for t, v in stats.items() if 'spconv' in t or 'linear' in t}
    print(json.dumps({'metric': 'training nodes/s (forward + backward + Adam step, DDP)', 'value': float(ms[1]) / (float(ms[0]) * 1e-3),
                      'unit': 'nodes/s', 'n_gpus': world, 'ms_per_step': float(ms[0]), 'scaling': 'weak', 'wgrad': args.wgrad,
                      'config': {'model': f'OccupancyNet C={args.channels} blocks={args.blocks}', 'nodes_per_rank': n_nodes, 'frames_per_rank': args.frames},
                      'loss_bits_per_node': [round(x, 4) for x in (losses[0], losses[-1])], 'kernels': k}))
if world > 1:
    dist.destroy_process_group()
