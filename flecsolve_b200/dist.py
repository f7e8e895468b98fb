"""Multi-process plumbing: one process per GPU, torch.distributed only for rendezvous,
broadcasting the NCCL unique id and max-over-ranks of timings.  The data path (halo exchange,
scalar all-reduces) is NCCL called from libfsb.so directly."""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


@dataclass
class World:
    rank: int
    local_rank: int
    size: int
    initialised: bool = False

    @property
    def is_root(self) -> bool:
        return self.rank == 0


def world_from_env(env=os.environ) -> World:
    return World(int(env.get("RANK", 0)), int(env.get("LOCAL_RANK", env.get("RANK", 0))), int(env.get("WORLD_SIZE", 1)))


def init(world: World, backend: str = "gloo") -> World:
    """Join the torch.distributed group described by MASTER_ADDR/MASTER_PORT (no-op for 1 rank)."""
    if world.size > 1 and not world.initialised:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if not dist.is_initialized():
            dist.init_process_group(backend=backend, rank=world.rank, world_size=world.size)
        world.initialised = True
    return world


def finalize(world: World) -> None:
    if world.size > 1 and world.initialised:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
        world.initialised = False


def broadcast_bytes(world: World, payload: bytes | None, src: int = 0) -> bytes:
    """Same bytes on every rank (used for the 128-byte NCCL unique id)."""
    if world.size == 1:
        assert payload is not None
        return payload
    import torch.distributed as dist
    box = [payload if world.rank == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def barrier(world: World) -> None:
    if world.size > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(world: World, value: float) -> float:
    if world.size == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(world: World, value: float) -> float:
    if world.size == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def equal_map(n: int, bins: int) -> np.ndarray:
    """flecsi::util::equal_map offsets: contiguous blocks, the first n % bins get one extra
    (the partition set_block_map installs, matrices/parcsr.hh:170-172)."""
    q, r = divmod(n, bins)
    sizes = np.full(bins, q, dtype=np.int64)
    sizes[:r] += 1
    out = np.zeros(bins + 1, dtype=np.int64)
    out[1:] = np.cumsum(sizes)
    return out


def make_context(world: World):
    """Create this rank's device context (device = LOCAL_RANK), sharing one NCCL clique."""
    from . import _lib as F
    uid = None
    if world.size > 1:
        uid = broadcast_bytes(world, F.Context.unique_id() if world.is_root else None)
    return F.Context(world.local_rank, world.rank, world.size, uid)
