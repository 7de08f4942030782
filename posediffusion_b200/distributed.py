"""Multi-GPU plumbing: the path shards embarrassingly over sequences.

One process per GPU (`torchrun`), contiguous blocks of sequences per rank, no communication inside the
sampling loop, and ONE collective at the end: an all-gather of the [B_local, N, 9] poses (NCCL over
NVLink on GPUs; gloo in the CPU tests).  Per-sequence RNG seeds are keyed by the GLOBAL sequence index so
results do not depend on the world size.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) block of `total` sequences owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sequence_seed(base_seed: int, global_index: int) -> int:
    return (base_seed * 1_000_003 + global_index * 7919 + 17) % (2**31 - 1)


def init_from_env(backend: str) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the process group if world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


_GATHER_BUFFERS: dict = {}


def gather_poses(local: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather the per-rank [B_local, N, 9] poses into [total, N, 9] in global sequence order.

    One `all_gather_into_tensor` on buffers that are allocated once per (shape, device) and reused: with equal blocks per
    rank (the benchmark configurations) the collective writes straight into the result and nothing else is launched; ragged
    blocks are padded to the widest one and compacted by one index_select."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(total, r, world) for r in range(world)]
    widest = max(hi - lo for lo, hi in sizes)
    tail = tuple(local.shape[1:])
    key = (local.device, local.dtype, widest, world, tail, total)
    bufs = _GATHER_BUFFERS.get(key)
    if bufs is None:
        out = torch.empty((world * widest,) + tail, device=local.device, dtype=local.dtype)
        pad = None if all(hi - lo == widest for lo, hi in sizes) else torch.zeros((widest,) + tail, device=local.device, dtype=local.dtype)
        keep = None
        if pad is not None:
            keep = torch.tensor([r * widest + i for r, (lo, hi) in enumerate(sizes) for i in range(hi - lo)], device=local.device)
        bufs = _GATHER_BUFFERS[key] = (out, pad, keep)
    out, pad, keep = bufs
    src = local.contiguous()
    if pad is not None:
        pad[: local.shape[0]] = local
        src = pad
    dist.all_gather_into_tensor(out, src)
    return out if keep is None else out.index_select(0, keep)
