"""Training slice on the GPU: the occupancy predictor trains through the tensor-core forward / dgrad /
wgrad kernels; the tensor-core wgrad and the per-offset library GEMMs give the same training trajectory; with two GPUs
DistributedDataParallel keeps the replicas identical (NCCL gradient all-reduce, reference train.py:139,215,359-404)."""
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(steps, wgrad_tc, seeds=(7001, 7002)):
    from fastpcc_b200 import autograd, train
    autograd.WGRAD_TC = wgrad_tc
    torch.manual_seed(0)
    model = train.OccupancyNet(64, 1).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    batch = train.make_batch(list(seeds), 5, torch.device('cuda'))
    losses = [float(train.train_step(model, opt, batch)) for _ in range(steps)]
    autograd.WGRAD_TC = True
    return losses, [p.detach().float().clone() for p in model.parameters()]


def test_occupancy_net_trains_and_wgrad_paths_agree():
    l_tc, p_tc = _run(6, True)
    l_lib, p_lib = _run(6, False)
    assert all(np.isfinite(l_tc)) and l_tc[-1] < l_tc[0] - 0.05, l_tc       # bits per node go down
    assert abs(l_tc[0] - l_lib[0]) < 1e-3 and abs(l_tc[-1] - l_lib[-1]) < 5e-2, (l_tc, l_lib)
    for a, b in zip(p_tc, p_lib):
        assert float((a - b).abs().max()) < 5e-2 * max(1.0, float(b.abs().max()))


def _ddp_worker(rank, world, port, out):
    import torch.distributed as dist
    from fastpcc_b200 import train
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(0)
    model = train.OccupancyNet(64, 1).to(dev)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank])
    opt = torch.optim.Adam(ddp.parameters(), lr=2e-3)
    batch = train.make_batch([7100 + rank], 5, dev)   # every rank its own scan
    losses = [float(train.train_step(ddp, opt, batch)) for _ in range(4)]
    flat = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        out.put((losses, float(max((g - gathered[0]).abs().max() for g in gathered))))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_ddp_replicas_stay_identical():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    losses, spread = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert spread == 0.0 and losses[-1] < losses[0], (losses, spread)
