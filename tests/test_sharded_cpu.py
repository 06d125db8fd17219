"""CPU (gloo, world_size 2): host-side logic of the tile-sharded multi-GPU path -- row partition, owner partition,
gradient packing and the reduce-to-owners exchange.  The per-rank partial gradients come from the C oracle rendering
only its band of tile rows is NOT needed here: the exchange is linear, so random partials test it exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from log_b200 import sharded


def test_tile_row_partition_covers_everything():
    for H in (16, 17, 1080, 2160, 33):
        gy = (H + 15) // 16
        for world in (1, 2, 3, 4, 8, 80):
            parts = sharded.tile_row_partition(H, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == gy
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_owner_partition_and_packing():
    for n in (0, 1, 7, 10_000_001):
        for world in (1, 2, 8):
            own = sharded.owner_partition(n, world)
            assert own[0][0] == 0 and own[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(own, own[1:]))
    n = 11
    g = (torch.randn(n, 3), torch.randn(n, 3), torch.randn(n), torch.randn(n, 3), torch.randn(n, 4), torch.randn(n, 3))
    buf = sharded.pack_grads(g)
    assert buf.shape == (n, sharded.GRAD_FLOATS_PRECOMP)
    for a, b in zip(g, sharded.unpack_grads(buf)):
        assert torch.equal(a, b)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    part = torch.randn(n, sharded.GRAD_FLOATS_PRECOMP, generator=g)
    mine = sharded.reduce_to_owners(part)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        out.put(torch.cat(gathered)[:n].numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize('n', [10, 1001])
def test_reduce_to_owners_gloo_world2(n):
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    [p.start() for p in procs]
    got = q.get(timeout=90)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    want = sum(torch.randn(n, sharded.GRAD_FLOATS_PRECOMP, generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
    np.testing.assert_allclose(got, want.numpy(), rtol=1e-6, atol=1e-6)
