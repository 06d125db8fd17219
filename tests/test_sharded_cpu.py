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


# ---- shard mode (SplatExchange) host logic -------------------------------------------------------------------------
def test_owner_of_row_inverts_tile_row_partition():
    from log_b200 import sharded
    for H in (16, 40, 90, 1080, 2160):
        for world in (1, 2, 3, 5, 8, 16, 32):
            for o, (a, b) in enumerate(sharded.tile_row_partition(H, world)):
                for y in range(a, b):
                    assert sharded.owner_of_row(y, H, world) == o, (H, world, y)


def test_shard_layout_regions_are_aligned_and_disjoint():
    from log_b200 import _capi, sharded
    for n, world in ((1, 1), (300, 4), (6001, 3), (10_000_000, 8)):
        lay, floats = sharded.shard_layout(n, world, world - 1)
        cap = sharded.owner_chunk(n, world)
        assert lay.cap == cap and lay.num_ranks == world and lay.my_rank == world - 1
        rows = world * cap
        regions = [(lay.off_count, world), (lay.off_splat, rows * 12), (lay.off_radii, rows), (lay.off_gid, rows),
                   (lay.off_dsplat, rows * 12), (lay.off_weight, rows), (lay.off_pcount, rows)]
        end = 0
        for off, size in regions:
            assert off % 64 == 0 and off >= end, (off, end)       # 256-byte aligned, in order, no overlap
            end = off + size
        assert floats >= end
        # every rank computes the same offsets (only my_rank differs)
        other, floats2 = sharded.shard_layout(n, world, 0)
        assert floats2 == floats and all(getattr(other, f) == getattr(lay, f) for f, _ in _capi.LgrShardLayout._fields_ if f != 'my_rank')
        # a shard never exceeds the rows one (source, owner) region can hold
        assert all(hi - lo <= cap for lo, hi in sharded.owner_partition(n, world))


def test_shard_send_scratch_size_matches_header_macro():
    import re
    from log_b200 import _capi
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'log_b200_raster.h')).read()
    assert re.search(r'#define LGR_SHARD_SEND_INTS\(n_local, r\) \(2 \* \(int64_t\)\(r\) \* \(\(\(\(n_local\) > 0 \? \(n_local\) : 1\) \+ 255\) / 256\) \+ \(r\)\)', hdr)
    for n, r in ((0, 4), (1, 1), (256, 2), (257, 2), (1_250_048, 8)):
        assert _capi.shard_send_ints(n, r) == 2 * r * ((max(n, 1) + 255) // 256) + r
    assert int(re.search(r'#define LGR_SHARD_MAX_RANKS (\d+)', hdr).group(1)) == _capi.LGR_SHARD_MAX_RANKS


def test_splat_exchange_rejects_cpu_buffers_and_bad_peer_lists():
    import pytest
    from log_b200 import sharded
    _, floats = sharded.shard_layout(1000, 2, 0)
    buf = torch.zeros(floats)
    from log_b200._capi import LgrError
    with pytest.raises(LgrError):
        sharded.SplatExchange(1000, 64, 0, 2, buf, [buf.data_ptr(), 0], barrier=lambda: None)      # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        sharded.SplatExchange(1000, 64, 0, 40, buf, [0] * 40, barrier=lambda: None)                # too many ranks
