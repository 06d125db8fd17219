"""Replacement for LoG/cuda/compute_radius.py (which JIT-compiles compute_radius_kernel.cu with glm).
Copy this file over LoG/cuda/compute_radius.py: `compute_radius_module.compute_radius(xyz, scaling, rotation,
proj_matrix, view_matrix, focal_x, focal_y, tanfovx, tanfovy)` keeps its signature
(call site LoG/model/level_of_gaussian.py:80-83)."""
import log_b200.rasterizer as compute_radius_module  # noqa: F401  (exposes .compute_radius)
