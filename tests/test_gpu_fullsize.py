"""GPU, BASELINE.json's full sizes.
  * config 1 (100k Gaussians, 1080p, SH degree 3): direct parity against the C oracle (it finishes in seconds).
  * 1M-Gaussian subset of config 3 at 1080p: direct parity against the fp32/fp64 C oracle.
  * config 3 (10M Gaussians, 1080p): size-independent properties -- invariance under a permutation of the Gaussians
    (the per-tile order is by (depth, index), so only exact depth ties could change anything), tile-row shards summing
    to the un-sharded result (the multi-GPU primitive), linearity of the backward in the cotangent, basic sanity."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, torch_dense as O
from util import f32_camera, record_parity, rel, run_gpu

pytestmark = pytest.mark.gpu
W, H = 1920, 1080


def f32_scene(sc):
    return {k: v.to(torch.float32).to(torch.float64) for k, v in sc.items()}


TOL = 1e-4      # north_star's bound; where the fp32 build of the oracle itself is further than that from its fp64 build
                # (config 1: sigma = 8 px splats), the bound is that fp32 floor -- see tests/test_gpu_parity.py:check_all


def compare(case, got, ref, ref32, keys):
    errs = {k: (rel(got[k], ref[k]), None if ref32 is None else rel(ref32[k], ref[k])) for k in keys}
    record_parity(case, errs, TOL, True)
    for k, (e, floor) in errs.items():
        assert e < max(TOL, 1.05 * (floor or 0.0)), (case, k, e, floor)


def test_config1_100k_sh3_direct_parity(built):
    cam = f32_camera(O.make_camera(W, H, sh_degree=3))
    sc = f32_scene(O.make_scene(100_000, W, H, 8.0, seed=0, sh_degree=3))
    sc.pop('colors')
    G = O.make_cotangent(3, H, W, seed=1).to(torch.float32).to(torch.float64)
    kw = dict(shs=sc['shs'], filter_mode=c_oracle.FILTER_ADD, dL_dimage=G)
    ref = c_oracle.render(cam, sc['means3D'], sc['opacities'], sc['scales'], sc['rotations'], dtype=np.float64, **kw)
    ref32 = c_oracle.render(cam, sc['means3D'], sc['opacities'], sc['scales'], sc['rotations'], dtype=np.float32, **kw)
    got = run_gpu(cam, sc, G, flavour='stock', sh_degree=3)
    compare('config1[100k,1080p,sh3,stock]', got, ref, ref32, ['image', 'dmeans3D', 'dmeans2D', 'dopacities', 'dscales', 'drotations', 'dshs'])
    rg = got['radii'].cpu().numpy()
    assert (rg != ref['radii']).sum() <= 20 and np.abs(rg - ref['radii']).max() <= 1


@pytest.fixture(scope='module')
def scene10m():
    cam = f32_camera(O.make_camera(W, H))
    sc = O.make_scene(10_000_000, W, H, 1.5, seed=0, dtype=torch.float32)
    G = O.make_cotangent(3, H, W, seed=1, dtype=torch.float32)
    return cam, sc, G


def test_config3_1m_subset_direct_parity(built, scene10m):
    cam, sc, G = scene10m
    sub = {k: v[:1_000_000].to(torch.float64) for k, v in sc.items()}
    kw = dict(colors_precomp=sub['colors'], filter_mode=c_oracle.FILTER_MAX, dL_dimage=G.to(torch.float64))
    ref = c_oracle.render(cam, sub['means3D'], sub['opacities'], sub['scales'], sub['rotations'], dtype=np.float64, **kw)
    ref32 = c_oracle.render(cam, sub['means3D'], sub['opacities'], sub['scales'], sub['rotations'], dtype=np.float32, **kw)
    got = run_gpu(cam, sub, G)
    compare('config3-subset[1M of 10M,1080p,fork]', got, ref, ref32, ['image', 'dmeans3D', 'dmeans2D', 'dopacities', 'dscales', 'drotations',
                                                                     'dcolors', 'point_weight', 'point_weight_pixel'])
    rg = got['radii'].cpu().numpy()
    assert (rg != ref['radii']).sum() <= 200 and np.abs(rg - ref['radii']).max() <= 1
    pid = got['point_id_pixel'].cpu().numpy()
    assert (pid != ref['point_id_pixel']).mean() < 1e-3


def test_config3_10m_direct_parity(built, scene10m):
    """The metric's own configuration (10 M Gaussians, 1080p, fork flavour, precomputed colour) against the fp64 C oracle,
    directly: image, all six gradient tensors and the aux outputs.  Bound: 1e-4, or the distance of the oracle's own fp32 build
    from its fp64 build where that is larger (it is, for this workload: sub-pixel splats at pixel coordinates ~1000 leave
    |d| ~ 1 px with ~1e-4 relative resolution in fp32 for ANY implementation) -- the CUDA path must be at least as accurate
    as a plain fp32 restatement.  The two oracle builds need ~1 minute on the box's host cores; they run once."""
    cam, sc, G = scene10m
    kw = dict(colors_precomp=sc['colors'], filter_mode=c_oracle.FILTER_MAX, dL_dimage=G.to(torch.float64))
    ref = c_oracle.render(cam, sc['means3D'], sc['opacities'], sc['scales'], sc['rotations'], dtype=np.float64, **kw)
    ref32 = c_oracle.render(cam, sc['means3D'], sc['opacities'], sc['scales'], sc['rotations'], dtype=np.float32, **kw)
    got = run_gpu(cam, sc, G)
    compare('config3[10M,1080p,fork]', got, ref, ref32, ['image', 'dmeans3D', 'dmeans2D', 'dopacities', 'dscales', 'drotations', 'dcolors',
                                                       'point_weight', 'point_weight_pixel'])
    rg = got['radii'].cpu().numpy()
    assert (rg != ref['radii']).sum() <= 2000 and np.abs(rg - ref['radii']).max() <= 1
    pid = got['point_id_pixel'].cpu().numpy()
    assert (pid != ref['point_id_pixel']).mean() < 1e-3


def test_config3_10m_permutation_invariance(built, scene10m):
    cam, sc, G = scene10m
    a = run_gpu(cam, sc, G)
    perm = torch.randperm(sc['means3D'].shape[0], generator=torch.Generator().manual_seed(5))
    b = run_gpu(cam, {k: v[perm] for k, v in sc.items()}, G)
    assert torch.isfinite(a['image']).all()
    # No float atomics touch the image, so it is bit-identical EXCEPT where two Gaussians of one tile share the exact
    # same fp32 depth (10M samples over ~27M representable depths in [2,20): ~1e6 equal pairs, a few hundred of them
    # in the same tile): their order is by index, which the permutation changes -- as it would in the reference.
    diff = (a['image'] - b['image']).detach().abs().amax(dim=0)
    assert float((diff > 1e-6).float().mean()) < 1e-3
    assert rel(b['image'], a['image']) < 1e-4
    p = perm.to(a['radii'].device)
    assert torch.equal(a['radii'][p], b['radii'])
    assert rel(b['point_weight'], a['point_weight'][p]) < 2e-3           # max is order independent (up to depth ties)
    assert float((a['point_weight'][p] != b['point_weight']).float().mean()) < 1e-3
    # gradients: identical up to fp32 atomics, except for the few Gaussians involved in an exact depth tie
    for k in ['dmeans3D', 'dmeans2D', 'dopacities', 'dscales', 'drotations', 'dcolors']:
        x, y = b[k].reshape(b[k].shape[0], -1), a[k][p].reshape(b[k].shape[0], -1)
        row_err = (x - y).norm(dim=1) / (y.norm(dim=1) + 1e-3 * y.norm(dim=1).mean())
        assert float((row_err > 1e-3).float().mean()) < 2e-3, k
        assert rel(x, y) < 2e-2, k
    pid_a, pid_b = a['point_id_pixel'], b['point_id_pixel']
    m = pid_a >= 0
    assert ((pid_b >= 0) == m).all()
    assert (p[pid_b[m].long()] == pid_a[m]).float().mean() > 0.999


def test_config3_10m_shards_sum_to_full_and_backward_is_linear(built, scene10m):
    cam, sc, G = scene10m
    full = run_gpu(cam, sc, G)
    assert int((full['radii'] > 0).sum()) == sc['means3D'].shape[0]
    assert float(full['image'].min()) >= 0.0 and float(full['image'].max()) < 1.5
    gy = (H + 15) // 16
    parts = [run_gpu(cam, sc, G, tile_rows=(r0, r1)) for r0, r1 in ((0, 17), (17, 34), (34, 51), (51, gy))]
    assert torch.equal(sum(p['image'] for p in parts).detach(), full['image'].detach())
    for k in ['dmeans3D', 'dmeans2D', 'dopacities', 'dscales', 'drotations', 'dcolors']:
        assert rel(sum(p[k] for p in parts), full[k]) < 2e-5, k
    g2 = run_gpu(cam, sc, -3.0 * G)
    for k in ['dmeans3D', 'dopacities', 'dscales', 'drotations', 'dcolors']:
        assert rel(g2[k], -3.0 * full[k]) < 1e-4, k      # fp32 atomics: accumulation order differs run to run
