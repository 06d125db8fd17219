"""Shard-mode checks shared by tests/test_zz_gpu_shard_mode.py (`-m gpu`, R virtual ranks on one GPU) and
tests/test_emulated_host.py (same code on the CPU SIMT emulation).  See either file for what is being compared."""
import numpy as np
import torch

from oracle import torch_dense as O
from util import device, f32_camera, rel, run_gpu, settings_from_camera


def f32_scene(sc):
    return {k: v.to(torch.float32).to(torch.float64) for k, v in sc.items()}


def sync():
    if device().type == 'cuda':
        torch.cuda.synchronize()


def make_ranks(n, H, world, dev):
    from log_b200 import sharded
    _, floats = sharded.shard_layout(n, world, 0)
    bufs = [torch.full((floats,), float('nan'), device=dev) for _ in range(world)]       # garbage where nothing is written
    ptrs = [b.data_ptr() for b in bufs]
    return [sharded.SplatExchange(n, H, r, world, bufs[r], ptrs, barrier=lambda: None) for r in range(world)]


def shard_step(ranks, settings, t, G, deg, flavour_filter):
    from log_b200 import sharded
    n = t['means3D'].shape[0]
    world = len(ranks)
    parts = sharded.owner_partition(n, world)
    steps = []
    for x, (lo, hi) in zip(ranks, parts):                          # phase 1 on every rank, then the "barrier"
        kw = dict(colors_precomp=t['colors'][lo:hi]) if deg == 0 else dict(shs=t['shs'][lo:hi])
        steps.append(x.project_and_send(settings, t['means3D'][lo:hi], t['opacities'][lo:hi], t['scales'][lo:hi],
                                        t['rotations'][lo:hi], filter_mode=flavour_filter, want_aux=True, **kw))
    outs = [x.receive_and_render(s) for x, s in zip(ranks, steps)]
    for x, s in zip(ranks, steps):
        x.blend_backward_and_return(s, G)
    back = [x.gather_and_project_backward(s) for x, s in zip(ranks, steps)]
    sync()
    return outs, back, steps


def compare(full, outs, back, deg, n_pix):
    image = sum(o[0] for o in outs)
    # bands are disjoint and zero elsewhere; the background of a band is composited by its owner only
    assert torch.equal(image, full['image'])
    assert torch.equal(torch.cat([o[1] for o in outs]), full['radii'])
    pid = torch.stack([o[2] for o in outs]).max(0).values
    assert torch.equal(pid, full['point_id_pixel'])
    assert torch.equal(sum(o[3] for o in outs), full['point_weight_pixel'])
    pw = torch.cat([b[1] for b in back])
    assert torch.equal(pw, full['point_weight'])
    pc = torch.cat([b[2] for b in back]).long()
    ids, cnt = torch.unique(full['point_id_pixel'], return_counts=True)
    ref_pc = torch.zeros_like(pc)
    ref_pc[ids[ids >= 0].long()] = cnt[ids >= 0]
    assert torch.equal(pc, ref_pc)
    names = ['dmeans3D', 'dmeans2D', 'dopacities', 'dscales', 'drotations'] + (['dcolors'] if deg == 0 else ['dshs'])
    idx = [0, 1, 2, 3, 4] + ([5] if deg == 0 else [6])
    for k, j in zip(names, idx):
        got = torch.cat([b[0][j] for b in back])
        assert torch.isfinite(got).all(), k
        assert rel(got, full[k]) < 2e-5, (k, rel(got, full[k]))


def run_two_steps(world, deg, flavour, size=(224, 160, 6001)):
    from log_b200._capi import LGR_FILTER_ADD, LGR_FILTER_MAX
    W, H, n = size
    dev = device()
    cam = f32_camera(O.make_camera(W, H, bg=(0.1, 0.2, 0.3), sh_degree=deg))
    ranks = make_ranks(n, H, world, dev)
    # two steps through the same buffers: the second has fewer visible rows (stale slots must be ignored) and other data
    for seed, spread in ((41, 4.0), (42, 9.0)):
        sc = f32_scene(O.make_scene(n, W, H, spread, seed=seed, sh_degree=deg))
        if deg > 0:
            sc.pop('colors', None)
        G = O.make_cotangent(3, H, W)
        full = run_gpu(cam, sc, G, flavour=flavour, sh_degree=deg)
        s = settings_from_camera(cam, dev, deg)
        t = {k: v.to(device=dev, dtype=torch.float32).contiguous() for k, v in sc.items()}
        t['opacities'] = t['opacities'].reshape(-1)
        Gd = G.to(device=dev, dtype=torch.float32)
        if flavour == 'stock':      # the stock flavour has no aux outputs through the public API: compare image + gradients
            outs, back, _ = shard_step(ranks, s, t, Gd, deg, LGR_FILTER_ADD)
            assert torch.equal(sum(o[0] for o in outs), full['image'])
            assert torch.equal(torch.cat([o[1] for o in outs]), full['radii'])
            for k, j in (('dmeans3D', 0), ('dmeans2D', 1), ('dopacities', 2), ('dscales', 3), ('drotations', 4), ('dshs', 6)):
                got = torch.cat([b[0][j] for b in back])
                assert rel(got, full[k]) < 2e-5, (k, rel(got, full[k]))
        else:
            outs, back, _ = shard_step(ranks, s, t, Gd, deg, LGR_FILTER_MAX)
            compare(full, outs, back, deg, W * H)


def run_device_sized_steps(world, size=(224, 160, 6001)):
    """`sync_free` (device-sized render: no read-back of D, CUDA-graph capturable): the first step is host-sized and learns
    the capacity; a second step on the same scene is device-sized and must give the same result; a third step with much
    bigger splats outgrows the capacity on some rank -- check_overflow() must say so -- and the redo (host-sized again, it
    re-learns the capacity) must be right."""
    from log_b200._capi import LGR_FILTER_MAX
    W, H, n = size
    dev = device()
    cam = f32_camera(O.make_camera(W, H, bg=(0.1, 0.2, 0.3)))
    ranks = make_ranks(n, H, world, dev)
    for x in ranks:
        x.sync_free = True
    s = settings_from_camera(cam, dev, 0)

    def scene(seed, spread):
        sc = f32_scene(O.make_scene(n, W, H, spread, seed=seed))
        G = O.make_cotangent(3, H, W)
        t = {k: v.to(device=dev, dtype=torch.float32).contiguous() for k, v in sc.items()}
        t['opacities'] = t['opacities'].reshape(-1)
        return run_gpu(cam, sc, G), t, G.to(device=dev, dtype=torch.float32)
    full, t, Gd = scene(41, 4.0)
    for k in range(2):      # host-sized, then device-sized
        outs, back, _ = shard_step(ranks, s, t, Gd, 0, LGR_FILTER_MAX)
        assert all(x.check_overflow()['overflow'] == 0 for x in ranks)
        assert all((x._inst_cap > 0) for x in ranks)
        compare(full, outs, back, 0, W * H)
    full, t, Gd = scene(42, 14.0)      # ~10 x the instances per band
    outs, back, _ = shard_step(ranks, s, t, Gd, 0, LGR_FILTER_MAX)
    flagged = 0
    for x in ranks:
        try:
            x.check_overflow()
        except RuntimeError:
            flagged += 1
    assert flagged > 0
    for x in ranks:      # a redo is host-sized on every rank (ranks that fitted are simply rendered again)
        x._inst_cap = 0
    outs, back, _ = shard_step(ranks, s, t, Gd, 0, LGR_FILTER_MAX)
    compare(full, outs, back, 0, W * H)


def run_empty_shards_and_bands():
    """More ranks than tile rows (empty bands) and fewer Gaussians than ranks*256 (empty shards)."""
    from log_b200._capi import LGR_FILTER_MAX
    W, H, n, world = 96, 40, 300, 4          # 3 tile rows for 4 ranks; owner chunk 256 -> ranks 2,3 own no Gaussians
    dev = device()
    cam = f32_camera(O.make_camera(W, H))
    sc = f32_scene(O.make_scene(n, W, H, 5.0, seed=7))
    G = O.make_cotangent(3, H, W)
    full = run_gpu(cam, sc, G)
    ranks = make_ranks(n, H, world, dev)
    t = {k: v.to(device=dev, dtype=torch.float32).contiguous() for k, v in sc.items()}
    t['opacities'] = t['opacities'].reshape(-1)
    outs, back, _ = shard_step(ranks, settings_from_camera(cam, dev), t, G.to(device=dev, dtype=torch.float32), 0, LGR_FILTER_MAX)
    compare(full, outs, back, 0, W * H)
