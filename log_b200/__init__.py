"""log_b200 -- B200-native (sm_100a) differentiable Gaussian-splatting rasteriser, drop-in for the hot path of
zju3dv/LoG (diff_gaussian_rasterization[_wodilate] + LoG/cuda).  See DESIGN.md / INTEGRATION.md."""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, StockGaussianRasterizer, compute_radius,
                         rasterize_forward, rasterize_backward, point_id_count)

__all__ = ['GaussianRasterizationSettings', 'GaussianRasterizer', 'StockGaussianRasterizer', 'compute_radius',
           'rasterize_forward', 'rasterize_backward', 'point_id_count']
