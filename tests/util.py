"""Shared helpers for the parity tests (oracle-side only utilities live in oracle/)."""
import numpy as np
import torch

from oracle import torch_dense as O


DEVICE = ['cuda:0']      # tests/conftest.py:emulated_backend switches this to 'cpu' (SIMT emulation of the kernels)


def device():
    return torch.device(DEVICE[0])


def rel(a, b):
    """Norm-wise relative error ||a-b|| / ||b||."""
    a = np.asarray(a.detach().cpu().numpy() if hasattr(a, 'detach') else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu().numpy() if hasattr(b, 'detach') else b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def record_parity(case, errors, bound, filter_on=True, note=None):
    """Append the achieved norm-wise relative errors of one parity case to gpurun_out/parity.json (hardware runs only; the
    file travels back with gpurun_out/ and is committed as profiles/parity_rNN.json).  errors: {tensor: (error, fp32-oracle
    error or None)}."""
    import json
    import os
    if DEVICE[0] == 'cpu':
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, 'gpurun_out', 'parity.json')
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[case] = {'bound': bound, 'low_pass_filter': bool(filter_on), 'note': note,
                      'errors': {k: {'vs_fp64_oracle': float(v[0]), 'fp32_oracle_vs_fp64_oracle': None if v[1] is None else float(v[1]),
                                     'exceeds_1e-4': bool(v[0] >= 1e-4), 'bound_used': max(bound, 1.05 * (v[1] or 0.0))} for k, v in errors.items()}}
        json.dump(data, open(path, 'w'), indent=1, sort_keys=True)
    except OSError:
        pass


def settings_from_camera(cam, device, sh_degree=None):
    from log_b200 import GaussianRasterizationSettings
    f = lambda t: t.to(device=device, dtype=torch.float32)
    return GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=f(cam.bg), scale_modifier=cam.scale_modifier, viewmatrix=f(cam.viewmatrix), projmatrix=f(cam.projmatrix),
        sh_degree=cam.sh_degree if sh_degree is None else sh_degree, campos=f(cam.campos), prefiltered=False, debug=False)


def f32_camera(cam):
    """Round every camera tensor to float32 (what the GPU path sees) but keep them as float64 for the fp64 oracle."""
    r = lambda t: t.to(torch.float32).to(torch.float64)
    return cam._replace(viewmatrix=r(cam.viewmatrix), projmatrix=r(cam.projmatrix), campos=r(cam.campos), bg=r(cam.bg),
                        tanfovx=float(np.float32(cam.tanfovx)), tanfovy=float(np.float32(cam.tanfovy)))


def run_gpu(cam, scene, G=None, flavour='fork', use_filter=True, sh_degree=0, device=None, tile_rows=None, cov3D=None):
    """Forward (+backward if G) through the public GaussianRasterizer API.  scene: dict of float tensors.
    cov3D (N,6): passed as the stock API's cov3D_precomp instead of scales / rotations; its gradient comes back as dcov3D."""
    from log_b200 import GaussianRasterizer, StockGaussianRasterizer
    dev = torch.device(device or DEVICE[0])
    s = settings_from_camera(cam, dev, sh_degree)
    rast = (GaussianRasterizer if flavour == 'fork' else StockGaussianRasterizer)(s)
    rast.tile_rows = tile_rows
    t = {k: v.to(device=dev, dtype=torch.float32).requires_grad_(G is not None) for k, v in scene.items()}
    n = t['means3D'].shape[0]
    m2d = torch.zeros(n, 3, device=dev, requires_grad=G is not None)
    kw = dict(means3D=t['means3D'], means2D=m2d, opacities=t['opacities'], scales=t['scales'], rotations=t['rotations'],
              cov3D_precomp=None)
    if cov3D is not None:
        c6 = cov3D.to(device=dev, dtype=torch.float32).requires_grad_(G is not None)
        kw.update(scales=None, rotations=None, cov3D_precomp=c6)
    if sh_degree > 0 or 'colors' not in t:
        kw.update(shs=t['shs'], colors_precomp=None)
    else:
        kw.update(shs=None, colors_precomp=t['colors'])
    if flavour == 'fork' and not use_filter:
        kw['use_filter'] = False
    out = rast(**kw)
    res = dict(image=out[0], radii=out[1])
    if flavour == 'fork':
        res.update(point_id_pixel=out[2], point_weight_pixel=out[3], point_weight=out[4])
    if G is not None:
        (out[0] * G.to(device=dev, dtype=torch.float32)).sum().backward()
        res.update(dmeans3D=t['means3D'].grad, dmeans2D=m2d.grad, dopacities=t['opacities'].grad.reshape(-1))
        if cov3D is not None:
            res.update(dcov3D=c6.grad)
        else:
            res.update(dscales=t['scales'].grad, drotations=t['rotations'].grad)
        if kw['shs'] is not None:
            res['dshs'] = t['shs'].grad
        else:
            res['dcolors'] = t['colors'].grad
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    return res
