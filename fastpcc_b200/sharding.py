"""Frame / partition sharding across the GPUs of one box.

Frames and kd-tree partitions are independent units (each becomes its own self-contained bitstream,
models/convolutional/lossl_coord_int/model.py:455-463), so the path shards with NO data-path collective: every
rank codes its own units; the only exchange is the host-side gather of the finished byte strings in original
order.  One process per GPU (`torch.distributed`, NCCL on the GPU box, gloo in the CPU tests).
"""
from typing import Callable, Dict, List, Optional, Sequence

import torch.distributed as dist


def assign(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Size-balanced deterministic assignment (longest-processing-time first); ties keep index order."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += int(sizes[i])
    for lst in out:
        lst.sort()
    return out


def local_indices(sizes: Sequence[int], rank: Optional[int] = None, world: Optional[int] = None, group=None) -> List[int]:
    """Unit indices of this rank.  `rank` / `world` default to the position in `group` (the default group if None)."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    return assign(sizes, world)[rank]


def gather_ordered(local: Dict[int, bytes], n_total: int, dst: int = 0, group=None) -> Optional[List[bytes]]:
    """Collects {unit index: bytes} from every rank on `dst`, returned in unit order (None elsewhere).
    `dst` is a rank WITHIN `group` (the default group if None); it is translated to the global rank gather_object takes."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        assert len(local) == n_total
        return [local[i] for i in range(n_total)]
    rank = dist.get_rank(group)
    bucket = [None] * dist.get_world_size(group) if rank == dst else None
    global_dst = dist.get_global_rank(group, dst) if group is not None else dst
    dist.gather_object(local, bucket, dst=global_dst, group=group)
    if rank != dst:
        return None
    merged: Dict[int, bytes] = {}
    for part in bucket:
        for k, v in part.items():
            assert k not in merged, f'unit {k} coded twice'
            merged[k] = v
    assert len(merged) == n_total, f'{n_total - len(merged)} units missing'
    return [merged[i] for i in range(n_total)]


def compress_sharded(compress_batch: Callable[[list], List[bytes]], units: list, sizes: Sequence[int],
                     dst: int = 0, group=None) -> Optional[List[bytes]]:
    """Every rank codes its share of `units` with `compress_batch` (e.g. Model.compress_batch); rank `dst` gets
    all bitstreams in the original order."""
    mine = local_indices(sizes, group=group)
    coded = compress_batch([units[i] for i in mine]) if mine else []
    return gather_ordered(dict(zip(mine, coded)), len(units), dst=dst, group=group)


def pack_partitions(streams: List[bytes]) -> bytes:
    """3-byte little-endian length prefix per stream (model.py:462)."""
    from . import bitstream
    return bitstream.pack_partitions(streams)


def unpack_partitions(blob: bytes) -> List[bytes]:
    from . import bitstream
    return bitstream.split_partitions(blob)
