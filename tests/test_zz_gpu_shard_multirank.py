"""GPU, world_size 2 over NCCL + torch symmetric memory (skipped on a single-GPU box): shard mode as it runs in production --
SplatExchange.over_symmetric_memory, forward()/backward() with their device-side barriers, through the autograd wrapper --
against the single-GPU result, two steps through the same buffers.  The same worker logic runs on the CPU as
tests/test_sharded_cpu.py::test_shard_mode_as_separate_processes_with_real_barriers (gloo, shared memory, emulated kernels).
Green on 2 B200s since round 2 (`gpurun --gpus 2 -- python -m pytest tests/test_zz_gpu_shard_multirank.py -m gpu`)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
W, H, N, STEPS = 320, 208, 20000, 2


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), here]
    from log_b200 import sharded
    from oracle import torch_dense as O
    from util import f32_camera, rel, run_gpu, settings_from_camera
    import util
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    util.DEVICE = [f'cuda:{rank}']
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    ok, detail = True, []
    try:
        xch = sharded.SplatExchange.over_symmetric_memory(N, H)
        cam = f32_camera(O.make_camera(W, H, bg=(0.1, 0.2, 0.3)))
        for step in range(STEPS):
            sc = O.make_scene(N, W, H, 5.0 if step % 2 == 0 else 1.5, seed=30 + step)
            G = O.make_cotangent(3, H, W, seed=step).to(device=dev, dtype=torch.float32)
            full = run_gpu(cam, sc, G)
            t = {k: v[xch.lo:xch.hi].to(device=dev, dtype=torch.float32).clone().requires_grad_(True) for k, v in sc.items()}
            m2d = torch.zeros(xch.hi - xch.lo, 3, device=dev, requires_grad=True)
            image, radii, pid, pwp = xch.rasterize(settings_from_camera(cam, dev), t['means3D'], m2d, t['opacities'], t['scales'],
                                                   t['rotations'], colors_precomp=t['colors'])
            (image * G).sum().backward()
            torch.cuda.synchronize()
            a, b = xch.band[0] * 16, min(xch.band[1] * 16, H)
            checks = {'image': torch.equal(image[:, a:b], full['image'][:, a:b]),
                      'radii': torch.equal(radii, full['radii'][xch.lo:xch.hi]),
                      'pid': torch.equal(pid[a:b], full['point_id_pixel'][a:b]),
                      'point_weight': torch.equal(xch.last_point_weight, full['point_weight'][xch.lo:xch.hi])}
            for k, g in (('dmeans3D', t['means3D'].grad), ('dmeans2D', m2d.grad), ('dopacities', t['opacities'].grad.reshape(-1)),
                         ('dscales', t['scales'].grad), ('drotations', t['rotations'].grad), ('dcolors', t['colors'].grad)):
                checks[k] = rel(g, full[k][xch.lo:xch.hi]) < 2e-5
            bad = [k for k, v in checks.items() if not v]
            ok &= not bad
            detail.append((step, bad))
    except Exception as e:      # report instead of leaving the peer waiting for the queue
        ok, detail = False, repr(e)
    q.put((rank, bool(ok), detail))
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_rank_shard_mode_matches_single_gpu(built):
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res, waited = {}, 0
    while len(res) < world and waited < 300:
        try:
            r, ok, detail = q.get(timeout=2)
            res[r] = (ok, detail)
        except Exception:
            waited += 2
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    [p.join(30) for p in procs]
    [p.kill() for p in procs if p.is_alive()]
    assert len(res) == world and all(v[0] for v in res.values()), res
