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
    make_tree_golden()
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


def make_tree_golden():
    """Index lists returned by the reference's own TensorTree.traverse (LoG/model/tensor_tree.py:164-186), on trees built
    with the reference's own initialize / split / remove, with a model stub whose compute_radius is the reference's
    PyTorch twin of the CUDA kernel (level_of_gaussian.py:68-69 `if False:` branch -> geometry.compute_radius).  The twin
    has no NDC cull, so the scenes keep every point inside +-1.25 NDC (checked), and no radius within 1e-3 (relative) of
    the keep / descend threshold (checked), so that float32 kernels must reproduce the lists exactly."""
    import importlib.util
    torch.set_default_dtype(torch.float64)
    import LoG.dataset.base as ref_base
    import LoG.model.geometry as ref_geo
    spec = importlib.util.spec_from_file_location('ref_tensor_tree', os.path.join(REF, 'LoG/model/tensor_tree.py'))
    tt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tt)
    out = {}
    rng = np.random.default_rng(31)
    W, H, fx = 640, 360, 500.0
    K = np.array([[fx, 0, W / 2], [0, fx, H / 2], [0, 0, 1.0]])
    cam = ref_base.prepare_camera(dict(W=W, H=H, K=K, R=np.eye(3), T=np.zeros((3, 1)), center=np.zeros((3, 1))), scale=1, znear=0.01, zfar=100.0)
    cam_t = {k: (torch.from_numpy(np.asarray(v, dtype=np.float64)) if isinstance(v, np.ndarray) else v) for k, v in cam.items()}
    out['tree_cam_spec'] = np.array([W, H, fx, fx, W / 2, H / 2], dtype=np.float64)
    out['tree_world_view_transform'] = np.asarray(cam['world_view_transform'], dtype=np.float64)
    out['tree_full_proj_transform'] = np.asarray(cam['full_proj_transform'], dtype=np.float64)
    out['tree_FoV'] = np.array([cam['FoVx'], cam['FoVy']], dtype=np.float64)
    case = 0
    for max_child, n_root, rounds in ((2, 70, 4), (4, 40, 3), (3, 1, 5)):
        tree = tt.TensorTree(max_child=max_child, max_level=20)
        tree.initialize(torch.zeros(n_root, 3))
        tree.log_query = False

        def fresh(n, depth):      # points in front of the camera, inside +-1.2 NDC; deeper levels are smaller
            z = rng.uniform(1.0, 12.0, n)
            x = rng.uniform(-1.2, 1.2, n) * (W / (2 * fx)) * z
            y = rng.uniform(-1.2, 1.2, n) * (H / (2 * fx)) * z
            px = np.exp(rng.normal(np.log(24.0) - 1.1 * depth, 0.9, n))            # pixel radius ~ 3 sigma
            sig = (px / 3.0) * z / fx
            scal = np.log(sig[:, None] * rng.uniform(0.3, 1.0, (n, 3)))
            rot = rng.normal(size=(n, 4)) * rng.uniform(0.2, 3.0, (n, 1))           # un-normalised, as stored
            return np.stack([x, y, z], -1), scal, rot
        xyz, scal, rot = fresh(n_root, 0)
        for rd in range(rounds):
            leaves = torch.where(tree.is_leaf & (tree.depth == rd))[0]
            pick = leaves[torch.from_numpy(rng.random(len(leaves)) < (0.7 if n_root > 1 else 0.9 if rd else 1.0))]
            if len(pick) == 0:
                break
            tree.split(pick)
            a, b, c = fresh(len(pick) * max_child, rd + 1)
            xyz, scal, rot = np.concatenate([xyz, a]), np.concatenate([scal, b]), np.concatenate([rot, c])
        # holes in the child table: the reference's remove() compacts every per-point array, do the same to the params
        cand = torch.where(tree.is_leaf & (~tree.is_root))[0]
        rm = cand[torch.from_numpy(rng.random(len(cand)) < 0.15)]
        flag_keep = torch.ones(tree.num_points, dtype=torch.bool)
        flag_keep[rm] = False
        tree.remove(rm)
        xyz, scal, rot = xyz[flag_keep.numpy()], scal[flag_keep.numpy()], rot[flag_keep.numpy()]
        assert xyz.shape[0] == tree.num_points

        class Stub:      # LoG/model/level_of_gaussian.py:64-93 with the `if False:` (PyTorch twin) branch taken
            def compute_radius(self, index, camera, level=0):
                scaling = torch.exp(torch.from_numpy(scal)[index])
                rotation = torch.nn.functional.normalize(torch.from_numpy(rot)[index])
                r2d = ref_geo.compute_radius(torch.from_numpy(xyz)[index], scaling, rotation, camera)
                return scaling.max(dim=-1).values, r2d
        stub = Stub()
        for _ in range(20):      # move points whose radius sits within 0.2 % of a threshold used below
            all_r = ref_geo.compute_radius(torch.from_numpy(xyz), torch.exp(torch.from_numpy(scal)),
                                           torch.nn.functional.normalize(torch.from_numpy(rot)), cam_t).numpy()
            near = np.zeros(len(all_r), bool)
            for thr in (3.0, 1.0, 8.0):
                near |= np.abs(all_r / thr - 1.0) < 2e-3
            if not near.any():
                break
            scal[near] += 0.03
        P = np.asarray(cam['full_proj_transform'], dtype=np.float64)
        hom = xyz @ P[:3] + P[3]
        assert (np.abs(hom[:, :2] / (hom[:, 3:4] + 1e-7)) < 1.25).all()
        pre = f'tree{case}_'
        out[pre + 'node_index'] = tree.node_index.numpy().astype(np.int32)
        out[pre + 'tree'] = tree.tree.numpy().astype(np.int32)
        out[pre + 'depth'] = tree.depth.numpy().astype(np.int8)
        out[pre + 'max_child'] = np.array([max_child, tree.max_level])
        out[pre + 'xyz'], out[pre + 'scaling_raw'], out[pre + 'rotation_raw'] = xyz, scal, rot
        out[pre + 'radius'] = all_r
        roots_all = torch.where(tree.is_root)[0]
        q = 0
        for min_px in (3.0, 1.0, 8.0):
            assert (np.abs(all_r / min_px - 1.0) > 1e-3).all(), 'a radius sits on the threshold: change the seed'
            for max_depth in (1000, 2, 0):
                for roots in (roots_all, roots_all[torch.from_numpy(rng.random(len(roots_all)) < 0.6)].flip(0)):
                    tree.min_resolution_pixel = min_px
                    idx = tree.traverse(stub, roots.long(), cam_t, max_depth=max_depth)
                    out[pre + f'q{q}_args'] = np.array([min_px, max_depth], dtype=np.float64)
                    out[pre + f'q{q}_roots'] = roots.numpy().astype(np.int64)
                    out[pre + f'q{q}_index'] = idx.numpy().astype(np.int64)
                    q += 1
        out[pre + 'num_queries'] = np.array([q])
        print(pre, tree, 'queries', q, 'last result', len(idx))
        case += 1
    out['num_trees'] = np.array([case])
    np.savez_compressed(os.path.join(HERE, 'reference_tree.npz'), **out)
    print('wrote reference_tree.npz')


if __name__ == '__main__':
    main()
