"""GPU, world_size 2 over NCCL (skipped on a single-GPU box): tile-band rendering + packed-row all-to-all to owner ranks
equals the single-GPU gradients.  Run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
W, H, N = 320, 208, 20000


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
    from log_b200 import rasterize_backward, rasterize_forward, sharded
    from log_b200._capi import LGR_FILTER_MAX
    from oracle import torch_dense as O
    from util import settings_from_camera
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    cam = O.make_camera(W, H, bg=(0.1, 0.2, 0.3), dtype=torch.float32)
    sc = O.make_scene(N, W, H, 5.0, seed=3, dtype=torch.float32)
    G = O.make_cotangent(3, H, W, dtype=torch.float32).to(dev)
    s = settings_from_camera(cam, dev)
    t = {k: v.to(dev) for k, v in sc.items()}
    op = t['opacities'].reshape(-1)
    band = sharded.tile_row_partition(H, world)[rank]
    img, radii, pid, pwp, pw, st = rasterize_forward(s, t['means3D'], op, t['scales'], t['rotations'], t['colors'], None,
                                                     LGR_FILTER_MAX, True, band, num_owners=world)
    rows = rasterize_backward(st, G, t['means3D'], op, t['scales'], t['rotations'], t['colors'], None)
    shard = sharded.exchange_rows_to_owners(rows, st.band_counts_host, N)          # NCCL all-to-all route
    px = sharded.PeerExchange(N)                                                    # fused NVLink push route
    for _ in range(2):                                                              # twice: staging buffers are reused
        shard2 = px.backward(st, G, t['means3D'], op, t['scales'], t['rotations'], t['colors'])
    dist.all_reduce(img)                                   # bands are disjoint: the sum is the full image
    shards = [torch.zeros_like(shard) for _ in range(world)]
    dist.all_gather(shards, shard)
    shards2 = [torch.zeros_like(shard2) for _ in range(world)]
    dist.all_gather(shards2, shard2)
    if rank == 0:
        q.put((img.cpu().numpy(), torch.cat(shards)[:N, :19].cpu().numpy(), torch.cat(shards2)[:N, :19].cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_rank_band_rendering_matches_single_gpu(built):
    from log_b200 import sharded
    from oracle import torch_dense as O
    from util import run_gpu
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    img, grads_full, grads2_full = q.get(timeout=600)
    grads, grads2 = grads_full[:, :17], grads2_full[:, :17]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    cam = O.make_camera(W, H, bg=(0.1, 0.2, 0.3), dtype=torch.float32)
    sc = O.make_scene(N, W, H, 5.0, seed=3, dtype=torch.float32)
    G = O.make_cotangent(3, H, W, dtype=torch.float32)
    full = run_gpu(cam, sc, G)
    dense = sharded.pack_grads((full['dmeans3D'], full['dmeans2D'], full['dopacities'], full['dscales'], full['drotations'],
                                full['dcolors'])).cpu().numpy()
    np.testing.assert_array_equal(img, full['image'].detach().cpu().numpy())
    err_nccl = float(np.linalg.norm(grads - dense) / np.linalg.norm(dense))
    err_peer = float(np.linalg.norm(grads2 - dense) / np.linalg.norm(dense))
    assert err_nccl < 2e-5 and err_peer < 2e-5, (err_nccl, err_peer)
    radii = full['radii'].cpu().numpy()
    np.testing.assert_array_equal(grads_full[:, 18].astype(np.int32), radii)
    np.testing.assert_array_equal(grads2_full[:, 18].astype(np.int32), radii)
