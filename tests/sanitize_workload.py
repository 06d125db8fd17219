"""Small workload that touches every kernel of the path (ordinary fork / stock SH-3 calls, long clustered-depth lists
that drive the in-smem sort's recursion, band mode with packed rows, point_id_count); meant to be run under
compute-sanitizer on the GPU box:

    compute-sanitizer --tool memcheck  python tests/sanitize_workload.py
    compute-sanitizer --tool racecheck python tests/sanitize_workload.py

Test infrastructure (lives under tests/; not collected by pytest).  Results of the last run
are recorded in DESIGN.md section 3 ("Sanitizer").
"""
import os
import sys, torch
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [_ROOT, os.path.join(_ROOT, 'tests')]
from oracle import torch_dense as O
from util import run_gpu, settings_from_camera
from log_b200 import rasterize_forward, rasterize_backward, sharded, point_id_count, GaussianRasterizer
from log_b200._capi import LGR_FILTER_MAX
torch.manual_seed(0)
# 1. ordinary fork + stock/SH paths, odd image size
cam = O.make_camera(150, 90, bg=(0.1, 0.2, 0.3), dtype=torch.float32)
sc = O.make_scene(1500, 150, 90, 4.0, seed=1, dtype=torch.float32)
G = O.make_cotangent(3, 90, 150, dtype=torch.float32)
run_gpu(cam, sc, G)
cam3 = O.make_camera(150, 90, sh_degree=3, dtype=torch.float32)
sc3 = O.make_scene(800, 150, 90, 4.0, seed=2, sh_degree=3, dtype=torch.float32); sc3.pop('colors')
run_gpu(cam3, sc3, G, flavour='stock', sh_degree=3)
# 2. long lists with clustered depths (sort recursion) on a tiny image
cam2 = O.make_camera(32, 32, dtype=torch.float32)
sc2 = O.make_scene(5000, 32, 32, 8.0, seed=3, dtype=torch.float32)
sc2['means3D'][:, 2] = 5.0 + 0.001 * torch.randn(5000)     # clustered depths
sc2['opacities'][:] = 0.02
run_gpu(cam2, sc2, O.make_cotangent(3, 32, 32, dtype=torch.float32))
# 3. band mode + rows + point count
dev = torch.device('cuda:0')
s = settings_from_camera(cam, dev)
t = {k: v.to(dev) for k, v in sc.items()}
op = t['opacities'].reshape(-1)
for r, band in enumerate(sharded.tile_row_partition(90, 3)):
    img, radii, pid, pwp, pw, st = rasterize_forward(s, t['means3D'], op, t['scales'], t['rotations'], t['colors'], None, LGR_FILTER_MAX, True, band, num_owners=3)
    rows = rasterize_backward(st, G.to(dev), t['means3D'], op, t['scales'], t['rotations'], t['colors'], None)
    sharded.rows_to_shard(rows, 0, 1500)
point_id_count(st.point_count)
torch.cuda.synchronize(); print('sanit workload done')
