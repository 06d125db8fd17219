"""Drop-in module with the name LoG imports at LoG/render/renderer.py:1,104 (fork flavour: 5-tuple return,
max(cov, 0.3) filter, `use_filter=` kwarg, `.compute_radius()`).  Put `<repo>/dropin` and `<repo>` on PYTHONPATH."""
from log_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401
