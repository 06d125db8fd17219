"""Drop-in module with the name LoG imports at LoG/render/renderer.py:100 and apps/check_gui.py:19 (stock flavour:
`(image, radii)` and the +0.3 covariance dilation).  Put `<repo>/dropin` and `<repo>` on PYTHONPATH."""
from log_b200.rasterizer import GaussianRasterizationSettings  # noqa: F401
from log_b200.rasterizer import StockGaussianRasterizer as GaussianRasterizer  # noqa: F401
