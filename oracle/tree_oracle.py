"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's tree walk, used as the checker for
log_b200/csrc/lgr_tree.cu.  Only tests/ may import this.

Follows LoG/model/tensor_tree.py:132-186 (`_query_tree_torch`, `traverse`) statement by statement, with the radius of
`model.compute_radius` (LoG/model/level_of_gaussian.py:64-93) taken from the compute_radius oracle (oracle/c_oracle.py,
CUDA-kernel semantics: NDC cull at +-1.3 -> radius 0) after the activations of LoG/model/activation.py:7,18
(exp, F.normalize).  Pinned against index lists produced by running the reference's own TensorTree.traverse
(tests/golden/reference_tree.npz, made by tests/golden/make_golden.py)."""
import numpy as np

from . import c_oracle


def radius2d(cam, xyz, scaling_raw, rotation_raw, index, dtype=np.float64):
    """level_of_gaussian.py:64-93: gather, activate, compute_radius."""
    index = np.asarray(index, dtype=np.int64)
    if index.size == 0:
        return np.zeros(0, dtype=dtype)
    sc = np.exp(np.asarray(scaling_raw, dtype=np.float64)[index])
    rot = np.asarray(rotation_raw, dtype=np.float64)[index]
    rot = rot / np.maximum(np.linalg.norm(rot, axis=-1, keepdims=True), 1e-12)
    return c_oracle.compute_radius(cam, np.asarray(xyz, dtype=np.float64)[index], sc, rot, dtype=dtype)


def traverse(cam, node_index, tree, max_level, min_resolution_pixel, xyz, scaling_raw, rotation_raw, root_index, max_depth=1000,
             dtype=np.float64, return_radii=False):
    node_index, tree = np.asarray(node_index, dtype=np.int64), np.asarray(tree, dtype=np.int64)
    root_index = np.asarray(root_index, dtype=np.int64)
    radii_seen = []
    # tensor_tree.py:166-176  the roots
    r2d = radius2d(cam, xyz, scaling_raw, rotation_raw, root_index, dtype)
    radii_seen.append(r2d)
    keep = (r2d < min_resolution_pixel) | (node_index[root_index] == -1)
    out = [root_index[keep]]
    index = root_index[~keep]
    # tensor_tree.py:133-163  _query_tree_torch
    level = 1
    while True:
        if level > max_level or level > max_depth:
            out.append(index)
            break
        index_node = node_index[index]
        id_child = tree[index_node].reshape(-1)
        id_child = id_child[id_child != -1]
        r2d = radius2d(cam, xyz, scaling_raw, rotation_raw, id_child, dtype)
        radii_seen.append(r2d)
        keep = (r2d < min_resolution_pixel) | (node_index[id_child] == -1)
        out.append(id_child[keep])
        if (~keep).sum() == 0:
            break
        index = id_child[~keep]
        level += 1
    res = np.concatenate(out) if out else root_index[:0]
    return (res, np.concatenate(radii_seen)) if return_radii else res
