"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard units with no data-path collective and
rank 0 reassembles the bitstreams in original order."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastpcc_b200 import sharding


def test_assignment_is_balanced_and_complete():
    sizes = [120, 50, 80, 119, 30, 77, 5]
    parts = sharding.assign(sizes, 3)
    assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(sizes)
    assert sharding.assign(sizes, 1) == [list(range(len(sizes)))]
    assert sharding.assign([], 2) == [[], []]


def test_partition_container_roundtrip():
    streams = [b'', b'\x01', bytes(range(200)) * 3]
    blob = sharding.pack_partitions(streams)
    assert blob[:3] == b'\x00\x00\x00' and sharding.unpack_partitions(blob) == streams


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sizes = [100, 10, 60, 55, 7]
        units = [bytes([i]) * s for i, s in enumerate(sizes)]
        seen = []

        def fake_compress(batch):  # stands in for Model.compress_batch: one stream per unit, order preserved
            seen.extend(batch)
            return [b'C' + u[:3] + len(u).to_bytes(2, 'little') for u in batch]

        out = sharding.compress_sharded(fake_compress, units, sizes)
        mine = sharding.local_indices(sizes)
        assert [units[i] for i in mine] == seen
        # the same through an explicit group with the LAST group rank as destination: assignment and gather must agree
        # on the position inside the group, and `dst` is a group rank (advisor finding of round 1)
        grp = dist.new_group(ranks=[0, 1])
        seen.clear()
        out_g = sharding.compress_sharded(fake_compress, units, sizes, dst=1, group=grp)
        assert [units[i] for i in sharding.local_indices(sizes, group=grp)] == seen
        if dist.get_rank(grp) == 1:
            assert out_g is not None and [o[1:4] for o in out_g] == [u[:3] for u in units]
        else:
            assert out_g is None
        if rank == 0:
            q.put(out)
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_ranks_shard_and_gather_in_order():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=90)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sizes = [100, 10, 60, 55, 7]
    assert out == [b'C' + (bytes([i]) * s)[:3] + s.to_bytes(2, 'little') for i, s in enumerate(sizes)]
