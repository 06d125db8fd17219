"""CPU: the WHOLE path -- real kernel source of projection, binning, sort, blend, both backward kernels and the real
C entry points of log_b200/csrc/lgr_capi.cu -- executed on the SIMT emulation (tests/emu) through the same C ABI the GPU
library exports, and compared with the fp64 C oracle, driving the C ABI from plain numpy buffers (no torch tensors involved).  The host
classes, band mode and shard mode run on the same emulation in tests/test_emulated_host.py.

What this does and does not show.  It executes the kernels' logic (indexing, staging, barriers, warp collectives, the
moment reduction, the chain rule) with IEEE float32 arithmetic; ex2.approx / rcp.approx are replaced by exact exp2f and
division and atomics are sequential, so low-order bits differ from the GPU.  It is a logic check that needs no GPU; the
`-m gpu` tests remain the parity proof on hardware.  Tolerance: the GPU tests' bound (norm-wise 1e-4, or 4x the error of
the float32 C oracle where fp32 itself cannot reach 1e-4).
"""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
from oracle import c_oracle, torch_dense as O   # noqa: E402
from util import f32_camera, rel                 # noqa: E402

F = np.float32
vp = ctypes.c_void_p


@pytest.fixture(scope='module')
def lib():
    import build_emu
    from log_b200._capi import LgrShardLayout, LgrView
    L = ctypes.CDLL(build_emu.build())
    i32, i64 = ctypes.c_int32, ctypes.c_int64
    V, Y = ctypes.POINTER(LgrView), ctypes.POINTER(LgrShardLayout)
    L.lgr_forward_project.argtypes = [V, i64] + [vp] * 13
    L.lgr_forward_render.argtypes = [V, i64, i64, i32, i32] + [vp] * 16
    L.lgr_backward.argtypes = [V, i64, i64] + [vp] * 23 + [i32, i64, vp]
    L.lgr_shard_send.argtypes = [V, Y, i64, i64, vp, vp, vp, vp, vp]
    L.lgr_shard_recv_bin.argtypes = [V, Y, vp, vp, vp, vp, vp, vp]
    L.lgr_blend_backward.argtypes = [V, i64, i64, vp, vp, vp, vp, vp, vp, vp]
    L.lgr_shard_return_rows.argtypes = [Y, vp, i64, vp, i32, i64, vp, vp]
    L.lgr_shard_gather.argtypes = [V, Y, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.lgr_compute_radius.argtypes = [i64, vp, vp, vp, vp, vp] + [ctypes.c_float] * 4 + [vp, vp]
    return L


def P(a):
    return None if a is None or a.size == 0 else vp(a.ctypes.data)


def ok(rc, what):
    assert rc == 0, (what, rc)


def make_view(cam, filter_mode, want_aux, K, rows, keep):
    from log_b200._capi import LgrView
    arrs = [np.ascontiguousarray(t.numpy(), dtype=F) for t in (cam.viewmatrix, cam.projmatrix, cam.campos, cam.bg)]
    keep.extend(arrs)
    v = LgrView()
    v.image_height, v.image_width = cam.image_height, cam.image_width
    v.tanfovx, v.tanfovy, v.scale_modifier = cam.tanfovx, cam.tanfovy, cam.scale_modifier
    v.sh_degree, v.sh_coeffs, v.filter_mode, v.want_aux = cam.sh_degree, K, filter_mode, int(want_aux)
    v.tile_row_begin, v.tile_row_end = (0, 0) if rows is None else rows
    v.viewmatrix_d, v.projmatrix_d, v.campos_d, v.bg_d = (a.ctypes.data for a in arrs)
    return v


def np_scene(sc):
    t = {k: np.ascontiguousarray(v.numpy(), dtype=F) for k, v in sc.items()}
    t['opacities'] = t['opacities'].reshape(-1)
    return t


def emu_render(lib, cam, t, G, filter_mode, deg, rows=None):
    """numpy mirror of log_b200/rasterizer.py rasterize_forward + rasterize_backward on the emulated library."""
    keep = []
    n = t['means3D'].shape[0]
    col, shs = (t['colors'], None) if deg == 0 else (None, t['shs'])
    K = 0 if shs is None else shs.shape[1]
    v = make_view(cam, filter_mode, True, K, rows, keep)
    H, W = cam.image_height, cam.image_width
    gx, gy = (W + 15) // 16, (H + 15) // 16
    ntiles = gx * (gy if rows is None else rows[1] - rows[0])
    splat, radii = np.zeros((n, 12), F), np.zeros(n, np.int32)
    clamped = np.zeros(n, np.uint8)
    tile_start, cursor, meta = np.zeros(ntiles + 1, np.int32), np.zeros(33 * max(ntiles, 1), np.int32), np.zeros(8, np.int32)
    ok(lib.lgr_forward_project(ctypes.byref(v), n, P(t['means3D']), P(t['opacities']), P(t['scales']), P(t['rotations']), P(col),
                               P(shs), P(splat), P(radii), P(clamped) if shs is not None else None, P(tile_start), P(cursor),
                               P(meta), None), 'project')
    D, max_len, num_long = int(meta[0]), int(meta[1]), int(meta[5])
    key, val, tmp = np.zeros(max(D, 1), np.uint32), np.zeros(max(D, 1), np.uint32), np.zeros(2 * max(D, 1), np.uint32)
    sorted_ids = np.zeros(max(D, 1), np.int32)
    image, final_T, n_contrib = np.zeros((3, H, W), F), np.ones((H, W), F), np.zeros((H, W), np.int32)
    pid, pwp, pw, pc = np.full((H, W), -1, np.int32), np.zeros((H, W), F), np.zeros(max(n, 1), F), np.zeros(max(n, 1), np.int32)
    ok(lib.lgr_forward_render(ctypes.byref(v), n, D, max_len, num_long, P(splat), P(radii), P(tile_start), P(cursor), P(key), P(val),
                              P(tmp), P(sorted_ids), P(image), P(final_T), P(n_contrib), P(pid), P(pwp), P(pw), P(pc), None), 'render')
    out = dict(image=image, radii=radii, point_id_pixel=pid, point_weight_pixel=pwp, point_weight=pw[:n], point_count=pc[:n],
               n_instances=D)
    if G is not None:
        dsplat = np.zeros((max(n, 1), 12), F)
        g = dict(dmeans3D=np.zeros((n, 3), F), dmeans2D=np.zeros((n, 3), F), dopacities=np.zeros(n, F), dscales=np.zeros((n, 3), F),
                 drotations=np.zeros((n, 4), F), dcolors=np.zeros((n, 3), F) if deg == 0 else None,
                 dshs=np.zeros_like(shs) if deg > 0 else None)
        ok(lib.lgr_backward(ctypes.byref(v), n, D, P(t['means3D']), P(t['opacities']), P(t['scales']), P(t['rotations']), P(col), P(shs),
                            P(splat), P(radii), P(clamped) if shs is not None else None, P(tile_start), P(sorted_ids), P(image), P(G),
                            P(dsplat), P(g['dmeans3D']), P(g['dmeans2D']), P(g['dopacities']), P(g['dscales']), P(g['drotations']),
                            P(g['dcolors']), P(g['dshs']), None, None, 0, 0, None), 'backward')
        out.update({k: a for k, a in g.items() if a is not None})
    return out


GRADS = ['dmeans3D', 'dmeans2D', 'dopacities', 'dscales', 'drotations']


def check_against_oracle(got, cam, sc, G, fm, deg):
    kw = dict(colors_precomp=sc['colors']) if deg == 0 else dict(shs=sc['shs'])
    args = (cam, sc['means3D'], sc['opacities'], sc['scales'], sc['rotations'])
    ref = c_oracle.render(*args, filter_mode=fm, dL_dimage=G, dtype=np.float64, **kw)
    ref32 = c_oracle.render(*args, filter_mode=fm, dL_dimage=G, dtype=np.float32, **kw)
    tol = lambda k: max(1e-4, 4.0 * rel(ref32[k], ref[k]))
    assert rel(got['image'], ref['image']) < tol('image')
    assert (got['radii'] != ref['radii']).sum() <= 1
    for k in GRADS + (['dcolors'] if deg == 0 else ['dshs']):
        assert rel(got[k], ref[k]) < tol(k), (k, rel(got[k], ref[k]), tol(k))
    assert rel(got['point_weight'], ref['point_weight']) < tol('point_weight')
    assert rel(got['point_weight_pixel'], ref['point_weight_pixel']) < tol('point_weight_pixel')
    assert (got['point_id_pixel'] != ref['point_id_pixel']).sum() <= 3
    ids, cnt = np.unique(got['point_id_pixel'], return_counts=True)
    want = np.zeros_like(got['point_count'])
    want[ids[ids >= 0]] = cnt[ids >= 0]
    assert np.array_equal(got['point_count'], want)


@pytest.mark.parametrize('W,H,n,r,deg,fm', [
    (64, 48, 500, 3.0, 0, c_oracle.FILTER_MAX),       # fork flavour, precomputed colours
    (50, 35, 300, 5.0, 3, c_oracle.FILTER_ADD),       # stock flavour, SH degree 3, image not a multiple of the tile size
    (32, 32, 1200, 2.0, 0, c_oracle.FILTER_NONE),     # dense small splats: several staging batches per tile, early stop
])
def test_full_path_emulated_matches_oracle(lib, W, H, n, r, deg, fm):
    cam = f32_camera(O.make_camera(W, H, bg=(0.2, 0.1, 0.3), sh_degree=deg))
    sc = {k: v.to(torch.float32).to(torch.float64) for k, v in O.make_scene(n, W, H, r, seed=3 + n, sh_degree=deg).items()}
    if deg > 0:
        sc.pop('colors', None)
    G = O.make_cotangent(3, H, W).to(torch.float32)
    got = emu_render(lib, cam, np_scene(sc), np.ascontiguousarray(G.numpy()), fm, deg)
    check_against_oracle(got, cam, sc, G.to(torch.float64), fm, deg)


def test_tile_row_bands_emulated_sum_to_the_full_image(lib):
    W, H, n = 64, 80, 400
    cam = f32_camera(O.make_camera(W, H, bg=(0.0, 0.3, 0.1)))
    t = np_scene(O.make_scene(n, W, H, 4.0, seed=9))
    G = np.ascontiguousarray(O.make_cotangent(3, H, W).numpy(), dtype=F)
    full = emu_render(lib, cam, t, G, c_oracle.FILTER_MAX, 0)
    parts = [emu_render(lib, cam, t, G, c_oracle.FILTER_MAX, 0, rows=b) for b in ((0, 2), (2, 3), (3, 5))]
    assert np.array_equal(sum(p['image'] for p in parts), full['image'])
    for k in GRADS + ['dcolors']:
        assert rel(sum(p[k] for p in parts), full[k]) < 2e-5, k
