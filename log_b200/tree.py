"""Fused level-of-Gaussian tree traversal -- host side of `lgr_tree_traverse`.

Drop-in for ``TensorTree.traverse(model, root_index, camera, max_depth)`` (LoG/model/tensor_tree.py:164-186), which LoG
calls once per training step from ``LevelOfGaussian.prepare`` (LoG/model/level_of_gaussian.py:243).  The reference walks
the tree level by level in Python: per level a gather of xyz / scaling / rotation, the activations, a
``compute_radius`` launch (level_of_gaussian.py:64-93), several boolean-mask kernels and a ``.sum() == 0`` host
synchronisation (tensor_tree.py:142-161).  Here the whole walk is enqueued at once; the host reads the result size once."""
import ctypes

import torch

from . import _capi


def traverse(tree, model, root_index, camera, max_depth=1000):
    """Same arguments and same result as the reference's ``tree.traverse``:
    tree   : object with ``node_index`` (P,) int32, ``tree`` (M, max_child) int32, ``max_child``, ``max_level``,
             ``min_resolution_pixel`` (a ``TensorTree``)
    model  : object with RAW parameters ``xyz`` (P,3), ``scaling`` (P,3), ``rotation`` (P,4) and ``activation`` (only the
             'exp' scaling activation is fused)
    camera : the rasteriser (``.raster_settings`` read as in level_of_gaussian.py:73-78)
    Returns the int64 index tensor ``index_concat`` -- same elements in the same order."""
    lib = _capi.load()
    act = getattr(model, 'activation', None)
    if act is not None and getattr(act, 'scaling_activation', torch.exp) is not torch.exp:
        raise NotImplementedError("only the 'exp' scaling activation (LoG/model/activation.py:7) is fused")
    rs = camera.raster_settings
    dev = tree.node_index.device

    def f32(t, name):
        _capi.require_cuda(t, name)
        t = t.detach()
        if t.dtype != torch.float32:
            raise TypeError(f'{name} must be float32, got {t.dtype}')
        return t.contiguous()

    def i32(t, name):
        _capi.require_cuda(t, name)
        if t.dtype != torch.int32:
            raise TypeError(f'{name} must be int32 (as TensorTree registers it), got {t.dtype}')
        return t.contiguous()
    xyz, sc, rot = f32(model.xyz, 'xyz'), f32(model.scaling, 'scaling'), f32(model.rotation, 'rotation')
    node_index, table = i32(tree.node_index, 'node_index'), i32(tree.tree, 'tree')
    P_, V_ = f32(rs.projmatrix, 'projmatrix'), f32(rs.viewmatrix, 'viewmatrix')
    roots = root_index.to(device=dev, dtype=torch.int64).contiguous()
    num_points, num_nodes, C = int(node_index.shape[0]), int(table.shape[0]), int(tree.max_child)
    if xyz.shape[0] != num_points or sc.shape[0] != num_points or rot.shape[0] != num_points:
        raise ValueError('model parameters and tree.node_index disagree on the number of points')
    t = _capi.LgrTree()
    t.num_points, t.num_nodes, t.max_child, t.max_level = num_points, num_nodes, C, int(tree.max_level)
    t.node_index_d = node_index.data_ptr() if num_points else None
    t.tree_d = table.data_ptr() if num_nodes else None
    slots = max(int(roots.shape[0]), num_nodes * C)
    scratch = torch.empty((_capi.tree_scratch_ints(num_points, slots),), dtype=torch.int32, device=dev)
    out = torch.empty((max(num_points, 1),), dtype=torch.int64, device=dev)
    count = torch.empty((1,), dtype=torch.int64, device=dev)
    p = lambda x: None if x.numel() == 0 else ctypes.c_void_p(x.data_ptr())
    fx = rs.image_width / (2.0 * rs.tanfovx)
    fy = rs.image_height / (2.0 * rs.tanfovy)
    _capi.check(lib.lgr_tree_traverse(ctypes.byref(t), p(xyz), p(sc), p(rot), p(P_), p(V_), float(fx), float(fy), float(rs.tanfovx),
                                      float(rs.tanfovy), p(roots), int(roots.shape[0]), float(tree.min_resolution_pixel),
                                      int(min(max_depth, 1 << 30)), p(scratch), p(out), p(count), _capi.current_stream()),
                'lgr_tree_traverse')
    return out[:int(count.item())]      # the one host synchronisation of the walk
