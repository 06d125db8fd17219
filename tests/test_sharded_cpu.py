"""CPU (gloo, world_size 2): host-side logic of the tile-sharded multi-GPU path -- row partition, owner partition,
gradient packing and the reduce-to-owners exchange.  The per-rank partial gradients come from the C oracle rendering
only its band of tile rows is NOT needed here: the exchange is linear, so random partials test it exactly."""
import os
import socket
import sys

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


# ---- shard mode as real processes: gloo barriers, exchange buffers in POSIX shared memory, kernels on the CPU emulation ----
def _shard_worker(rank, world, port, tag, n, W, H, steps, out):
    """One rank of SplatExchange.forward()/backward() (through the autograd wrapper) with the real barrier structure."""
    import ctypes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, 'tests'), os.path.join(root, 'tests', 'emu')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import build_emu
    import util
    from log_b200 import _capi
    from oracle import torch_dense as O
    from util import f32_camera, rel, run_gpu, settings_from_camera
    _capi._lib = _capi.bind(ctypes.CDLL(build_emu.build()))          # TEST ONLY: the emulated kernels (tests/emu)
    _capi.current_stream = lambda device=None: None
    _capi.require_cuda = lambda t, name: None
    util.DEVICE = ['cpu']
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    _, floats = sharded.shard_layout(n, world, rank)
    maps = [torch.from_file(f'/dev/shm/lgr_xch_{tag}_{r}', shared=True, size=floats, dtype=torch.float32) for r in range(world)]
    if True:      # every rank initialises its own buffer, then everybody waits
        maps[rank].fill_(float('nan'))
    dist.barrier()
    xch = sharded.SplatExchange(n, H, rank, world, maps[rank], [m.data_ptr() for m in maps], barrier=dist.barrier)
    dist.barrier()
    cam = f32_camera(O.make_camera(W, H, bg=(0.1, 0.2, 0.3)))
    ok = True
    for step in range(steps):
        sc = O.make_scene(n, W, H, 4.0 if step % 2 == 0 else 1.5, seed=70 + step)
        G = O.make_cotangent(3, H, W, seed=step).to(torch.float32)
        full = run_gpu(cam, sc, G)                                     # single-rank reference, computed by every process
        t = {k: v[xch.lo:xch.hi].to(torch.float32).clone().requires_grad_(True) for k, v in sc.items()}
        m2d = torch.zeros(xch.hi - xch.lo, 3, requires_grad=True)
        image, radii, pid, pwp = xch.rasterize(settings_from_camera(cam, torch.device('cpu')), t['means3D'], m2d, t['opacities'],
                                               t['scales'], t['rotations'], colors_precomp=t['colors'])
        (image * G).sum().backward()                                    # loss on this rank's band only (zero elsewhere)
        a, b = xch.band[0] * 16, min(xch.band[1] * 16, H)
        ok &= torch.equal(image[:, a:b], full['image'][:, a:b]) and float(image[:, :a].abs().sum() + image[:, b:].abs().sum()) == 0.0
        ok &= torch.equal(radii, full['radii'][xch.lo:xch.hi]) and torch.equal(pid[a:b], full['point_id_pixel'][a:b])
        ok &= torch.equal(xch.last_point_weight, full['point_weight'][xch.lo:xch.hi])
        for k, g in (('dmeans3D', t['means3D'].grad), ('dmeans2D', m2d.grad), ('dopacities', t['opacities'].grad.reshape(-1)),
                     ('dscales', t['scales'].grad), ('drotations', t['rotations'].grad), ('dcolors', t['colors'].grad)):
            want = full[k][xch.lo:xch.hi]
            ok &= bool(torch.isfinite(g).all()) and (rel(g, want) < 2e-5 or float(want.abs().max()) == 0.0)
    # evaluation: two forwards in a row without a backward (the entry barrier of forward() is needed here)
    for step in range(2):
        sc = O.make_scene(n, W, H, 3.0, seed=90 + step)
        full = run_gpu(cam, sc, None)
        t = {k: v[xch.lo:xch.hi].to(torch.float32) for k, v in sc.items()}
        image, radii, pid, pwp, _ = xch.forward(settings_from_camera(cam, torch.device('cpu')), t['means3D'], t['opacities'].reshape(-1),
                                                t['scales'], t['rotations'], colors_precomp=t['colors'])
        a, b = xch.band[0] * 16, min(xch.band[1] * 16, H)
        ok &= torch.equal(image[:, a:b], full['image'][:, a:b]) and torch.equal(pid[a:b], full['point_id_pixel'][a:b])
    out.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_shard_mode_as_separate_processes_with_real_barriers(world):
    """What the single-process emulation cannot exercise: SplatExchange.forward()/backward() as written, i.e. with their
    barriers, run by `world` concurrent processes (gloo) whose exchange buffers live in POSIX shared memory mapped at
    different addresses in every process -- exactly how NVLink peer mappings are addressed.  Two steps through the same
    buffers; every rank checks its band and its own Gaussians' gradients against the single-rank result."""
    import uuid
    n, W, H, steps = 500, 64, 80, 2
    emu_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu')
    if emu_dir not in sys.path:
        sys.path.insert(0, emu_dir)
    import build_emu
    build_emu.build()                      # once, here: the workers must not race to rebuild it
    tag = uuid.uuid4().hex[:8]
    _, floats = sharded.shard_layout(n, world, 0)
    files = [f'/dev/shm/lgr_xch_{tag}_{r}' for r in range(world)]
    for f in files:
        with open(f, 'wb') as fh:
            fh.truncate(floats * 4)
    try:
        port = _free_port()
        ctx = mp.get_context('spawn')
        q = ctx.Queue()
        procs = [ctx.Process(target=_shard_worker, args=(r, world, port, tag, n, W, H, steps, q)) for r in range(world)]
        [p.start() for p in procs]
        res, waited = {}, 0.0
        while len(res) < world and waited < 240:
            try:
                r, ok = q.get(timeout=2)
                res[r] = ok
            except Exception:
                waited += 2
                if any(p.exitcode not in (None, 0) for p in procs):      # a worker died: do not wait for the others
                    break
        [p.join(30) for p in procs]
        [p.kill() for p in procs if p.is_alive()]
        assert all(p.exitcode == 0 for p in procs)
        assert res == {r: True for r in range(world)}, res
    finally:
        for f in files:
            if os.path.exists(f):
                os.remove(f)
