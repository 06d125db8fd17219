"""Generate golden vectors FROM THE REFERENCE ITSELF (run in the build container only; /root/reference does not
exist on the GPU box).  Usage:  python tests/golden/make_golden.py

What is pinned (everything of the hot path that exists in-tree in the reference and runs on CPU):
  * camera matrices            LoG/dataset/base.py:20-55 prepare_camera, LoG/utils/camera.py:7-29
  * 3D covariance              LoG/model/geometry.py:27-41 computeCov3D (+ build_rotation :4-25)
  * EWA 2D covariance          LoG/model/geometry.py:91-130 computeCov2D0 (clamp_min 0.3 variant)
  * projected 3-sigma radius   LoG/model/geometry.py:132-151 compute_radius
  * SH colour                  LoG/model/sh_utils.py:31-73 eval_sh_wobase / SH2RGB  (as used by activation.py:27-34)
The blend itself is NOT in the reference tree (external un-vendored CUDA packages), so no golden exists for it.
"""
import os
import sys

import numpy as np
import torch

REF = '/root/reference'
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.set_default_dtype(torch.float64)          # reference allocates J with the default dtype (geometry.py:111)
    import LoG.dataset.base as ref_base
    import LoG.model.geometry as ref_geo
    import LoG.model.sh_utils as ref_sh

    rng = np.random.default_rng(20260922)
    out = {}
    cams = []
    # three cameras: identity, rotated+translated, off-centre principal point
    def rot(ax, ang):
        ax = np.asarray(ax, dtype=np.float64); ax /= np.linalg.norm(ax)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    specs = [
        dict(W=640, H=360, fx=554.0, fy=554.0, cx=320.0, cy=180.0, R=np.eye(3), T=np.zeros((3, 1))),
        dict(W=320, H=240, fx=300.0, fy=310.0, cx=160.0, cy=120.0, R=rot([0.2, 1.0, 0.1], 0.4), T=np.array([[0.3], [-0.2], [1.5]])),
        dict(W=256, H=256, fx=200.0, fy=200.0, cx=140.0, cy=110.0, R=rot([1.0, 0.3, -0.5], -0.7), T=np.array([[-0.5], [0.4], [2.0]])),
    ]
    for ci, s in enumerate(specs):
        K = np.array([[s['fx'], 0, s['cx']], [0, s['fy'], s['cy']], [0, 0, 1.0]])
        center = -(s['R'].T @ s['T'])
        cam_in = dict(W=s['W'], H=s['H'], K=K, R=s['R'], T=s['T'], center=center)
        cam = ref_base.prepare_camera(cam_in, scale=1, znear=0.01, zfar=100.0)
        N = 256
        # points in front of the camera, some far off-axis (exercise the 1.3*tanfov clamp), in WORLD coords
        z = rng.uniform(0.5, 20.0, N)
        nx = rng.uniform(-1.6, 1.6, N) * (s['W'] / (2 * s['fx']))
        ny = rng.uniform(-1.6, 1.6, N) * (s['H'] / (2 * s['fy']))
        pc = np.stack([nx * z, ny * z, z], -1)
        xyz = (pc - s['T'].reshape(1, 3)) @ s['R']              # world = R^T (cam - T)
        scaling = np.exp(rng.normal(-2.5, 1.0, (N, 3)))
        rotq = rng.normal(size=(N, 4)); rotq /= np.linalg.norm(rotq, axis=-1, keepdims=True)
        cam_t = {k: (torch.from_numpy(np.asarray(v, dtype=np.float64)) if isinstance(v, np.ndarray) else v) for k, v in cam.items()}
        xyz_t, sc_t, rq_t = map(lambda a: torch.from_numpy(a), (xyz, scaling, rotq))
        cov3 = ref_geo.computeCov3D(sc_t, rq_t)
        a, b, c = ref_geo.computeCov2D0(cov3, xyz_t, cam_t['world_view_transform'], cam_t)
        rad = ref_geo.compute_radius(xyz_t, sc_t, rq_t, cam_t)
        pre = f'cam{ci}_'
        out[pre + 'spec'] = np.array([s['W'], s['H'], s['fx'], s['fy'], s['cx'], s['cy']], dtype=np.float64)
        out[pre + 'R'] = s['R']; out[pre + 'T'] = s['T']
        out[pre + 'world_view_transform'] = np.asarray(cam['world_view_transform'], dtype=np.float64)
        out[pre + 'full_proj_transform'] = np.asarray(cam['full_proj_transform'], dtype=np.float64)
        out[pre + 'projection_matrix'] = np.asarray(cam['projection_matrix'], dtype=np.float64)
        out[pre + 'camera_center'] = np.asarray(cam['camera_center'], dtype=np.float64)
        out[pre + 'FoV'] = np.array([cam['FoVx'], cam['FoVy']], dtype=np.float64)
        out[pre + 'xyz'] = xyz; out[pre + 'scaling'] = scaling; out[pre + 'rotation'] = rotq
        out[pre + 'cov3D'] = cov3.numpy()
        out[pre + 'cov2D'] = np.stack([a.numpy(), b.numpy(), c.numpy()], -1)
        out[pre + 'radius'] = rad.numpy()
    # SH
    N = 128
    dirs = rng.normal(size=(N, 3)); dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    dc = rng.normal(size=(N, 3))
    rest = rng.normal(size=(N, 15, 3)) * 0.3
    out['sh_dirs'] = dirs; out['sh_dc'] = dc; out['sh_rest'] = rest
    for deg in (1, 2, 3):
        K = (deg + 1) ** 2 - 1
        val = ref_sh.SH2RGB(torch.from_numpy(dc)) + ref_sh.eval_sh_wobase(torch.from_numpy(dirs), torch.from_numpy(rest[:, :K]), degree=deg)
        out[f'sh_rgb_deg{deg}'] = val.numpy()
    out['sh_rgb_deg0'] = ref_sh.SH2RGB(torch.from_numpy(dc)).numpy()
    # sparse Adam (LoG/model/sparse_optimizer.py:41-78 _single_tensor_adam), float32 as LoG runs it
    make_adam_golden()
    np.savez_compressed(os.path.join(HERE, 'reference_geometry_sh.npz'), **out)
    print('wrote', os.path.join(HERE, 'reference_geometry_sh.npz'), {k: v.shape for k, v in out.items() if 'cam0' in k or 'sh' in k})


def make_adam_golden():
    torch.set_default_dtype(torch.float32)
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_sparse_optimizer', os.path.join(REF, 'LoG/model/sparse_optimizer.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(77)
    N, K = 500, 180
    out = {}
    for name, C in (('xyz', 3), ('rotation', 4), ('opacity', 1)):
        param = torch.randn(N, C, generator=g)
        m = torch.randn(N, C, generator=g) * 0.01
        v = torch.rand(N, C, generator=g) * 1e-4
        vmax = v * (1 + torch.rand(N, C, generator=g))
        index = torch.randperm(N, generator=g)[:K].sort().values
        for step, lr, ams in ((1, 1.6e-4, False), (37, 5e-3, False), (1000, 1e-3, True)):
            grad = torch.randn(K, C, generator=g) * (10.0 ** float(torch.randint(-4, 1, (1,), generator=g)))
            p_in, m_in, v_in, vm_in = param.clone(), m.clone(), v.clone(), vmax.clone()
            # what SparseOptimizer.step does around the call (sparse_optimizer.py:163-196): gather, update, scatter back
            ps, ms, vs = p_in[index].clone(), m_in[index].clone(), v_in[index].clone()
            vms = vm_in[index].clone() if ams else None
            ps, ms, vs, vms = mod._single_tensor_adam(ps, grad, ms, vs, vms, step, lr, eps=1e-15)
            p_out, m_out, v_out, vm_out = p_in.clone(), m_in.clone(), v_in.clone(), vm_in.clone()
            p_out[index], m_out[index], v_out[index] = ps, ms, vs
            if ams:
                vm_out[index] = vms
            key = f'adam_{name}_s{step}_'
            for k_, t_ in (('param_in', p_in), ('m_in', m_in), ('v_in', v_in), ('vmax_in', vm_in), ('grad', grad),
                           ('param_out', p_out), ('m_out', m_out), ('v_out', v_out), ('vmax_out', vm_out)):
                out[key + k_] = t_.numpy()
            out[key + 'index'] = index.numpy().astype(np.int64)
            out[key + 'hyper'] = np.array([step, lr, 0.9, 0.999, 1e-15, float(ams)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'reference_sparse_adam.npz'), **out)
    print('wrote reference_sparse_adam.npz', len(out))
    torch.set_default_dtype(torch.float64)


if __name__ == '__main__':
    main()
