"""Data-parallel plumbing: graphs are independent units, so rank r owns a contiguous shard of the
graphs (its own block-diagonal CSR and features) and the ONLY exchange per optimiser step is one
all-reduce(sum) of the flattened weight gradients [dW1|db1|dW2|db2] (502 003 floats = 2.0 MB at
F=1000,H=500,K=3) plus the loss scalar -- NCCL over NVLink on GPUs, gloo in the CPU tests.

The reference has no distributed code at all; this is new capability layered on
the per-graph independence of its training loop (python/Training/TrainingNeural.py:371-388).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed when WORLD_SIZE > 1; one process per GPU."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `total` graphs for `rank`; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_reduce_max_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def flat_grad_layout(sizes) -> Tuple[list, int]:
    """Offsets of 16-byte aligned segments inside the flat gradient buffer (mirrors GCNEngine)."""
    offs, tot = [], 0
    for s in sizes:
        offs.append(tot)
        tot += (int(s) + 3) & ~3
    return offs, tot
