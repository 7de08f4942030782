"""world_size-2 gloo test of the sequence sharding + final gather (the path's only collective)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from posediffusion_b200.distributed import gather_poses, sequence_seed, shard_range


def test_shard_range_partitions():
    for total in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(h - l for l, h in blocks) - min(h - l for l, h in blocks) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
    assert len({sequence_seed(0, i) for i in range(64)}) == 64


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    local = torch.stack([torch.full((3, 9), float(i)) for i in range(lo, hi)]) if hi > lo else torch.zeros(0, 3, 9)
    full = gather_poses(local, total)
    q.put((rank, full[:, 0, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 8])
def test_gather_poses_world2_gloo(total):
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert got[r] == [float(i) for i in range(total)]
