"""CPU: the C restatement (tile based, analytic backward) equals the dense autograd definition."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, torch_dense as O
from util import rel

CASES = [
    # W, H, n, radius px, sh degree, filter
    (64, 48, 300, 3.0, 0, O.FILTER_ADD),
    (80, 64, 300, 4.0, 3, O.FILTER_MAX),
    (70, 50, 250, 2.0, 2, O.FILTER_NONE),
    (50, 35, 200, 8.0, 1, O.FILTER_MAX),
]


@pytest.mark.parametrize('W,H,n,r,deg,fm', CASES)
def test_c_oracle_matches_autograd_definition(built, W, H, n, r, deg, fm):
    cam = O.make_camera(W, H, bg=(0.2, 0.5, 0.7), sh_degree=deg,
                        R=[[0.98, 0.0, 0.199], [0, 1, 0], [-0.199, 0, 0.98]], T=[0.1, -0.05, 0.3])
    sc = O.make_scene(n, W, H, r, sh_degree=deg, seed=3)
    sc['means3D'][:5, 2] = -1.0          # behind the camera
    sc['means3D'][5:10, 0] += 50         # far off screen
    sc['opacities'][10:14] = 0.001       # below 1/255 everywhere
    sc['opacities'][14:18] = 1.0         # alpha clamp 0.99 active (straight-through in backward)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    m2d = torch.zeros(n, 3, dtype=torch.float64, requires_grad=True)
    kw = dict(colors_precomp=leaves['colors']) if deg == 0 else dict(shs=leaves['shs'])
    out = O.render(leaves['means3D'], leaves['opacities'], leaves['scales'], leaves['rotations'], cam, means2D=m2d,
                   filter_mode=fm, **kw)
    G = O.make_cotangent(3, H, W)
    (out['image'] * G).sum().backward()
    kw2 = dict(colors_precomp=sc['colors']) if deg == 0 else dict(shs=sc['shs'])
    co = c_oracle.render(cam, sc['means3D'], sc['opacities'], sc['scales'], sc['rotations'], filter_mode=fm, dL_dimage=G,
                         dtype=np.float64, **kw2)
    assert co['n_instances'] == out['n_instances']
    assert rel(co['image'], out['image']) < 1e-12
    np.testing.assert_array_equal(co['radii'], out['radii'].numpy())
    np.testing.assert_array_equal(co['point_id_pixel'], out['point_id_pixel'].numpy())
    assert rel(co['point_weight_pixel'], out['point_weight_pixel']) < 1e-12
    assert rel(co['point_weight'], out['point_weight']) < 1e-12
    assert rel(co['final_T'], out['final_T']) < 1e-12
    assert rel(co['dmeans3D'], leaves['means3D'].grad) < 1e-10
    assert rel(co['dmeans2D'], m2d.grad) < 1e-10
    assert rel(co['dopacities'], leaves['opacities'].grad.reshape(-1)) < 1e-10
    assert rel(co['dscales'], leaves['scales'].grad) < 1e-10
    assert rel(co['drotations'], leaves['rotations'].grad) < 1e-10
    if deg == 0:
        assert rel(co['dcolors'], leaves['colors'].grad) < 1e-10
    else:
        assert rel(co['dshs'], leaves['shs'].grad) < 1e-10
    # float32 build of the same C code stays within fp32 round-off of it
    c32 = c_oracle.render(cam, sc['means3D'], sc['opacities'], sc['scales'], sc['rotations'], filter_mode=fm, dL_dimage=G,
                          dtype=np.float32, **kw2)
    assert rel(c32['image'], co['image']) < 1e-5
    assert rel(c32['dmeans3D'], co['dmeans3D']) < 1e-4


def test_empty_and_all_culled(built):
    cam = O.make_camera(32, 32, bg=(0.3, 0.6, 0.9))
    e = np.zeros((0, 3))
    out = c_oracle.render(cam, e, np.zeros(0), e, np.zeros((0, 4)), colors_precomp=e)
    assert out['n_instances'] == 0
    np.testing.assert_allclose(out['image'], np.broadcast_to(np.array([0.3, 0.6, 0.9])[:, None, None], (3, 32, 32)))
    assert (out['point_id_pixel'] == -1).all()
    sc = O.make_scene(10, 32, 32, 3.0)
    sc['means3D'][:, 2] = -5.0
    out = c_oracle.render(cam, sc['means3D'], sc['opacities'], sc['scales'], sc['rotations'], colors_precomp=sc['colors'])
    assert out['n_instances'] == 0 and (out['radii'] == 0).all()
