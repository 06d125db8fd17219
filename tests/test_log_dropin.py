"""CPU, build container only (skipped where /root/reference is absent, e.g. on the GPU box): LoG's own
`LoG/render/renderer.py` imports and runs UNMODIFIED with `dropin/` on the path, binds to this repo's classes for
both flavours (renderer.py:1, 99-105), builds the rasteriser through `BaseRender.prepare` (:57-78), and its
`render()` (:117-153) reaches our `GaussianRasterizer.forward` with exactly the keyword set it passes."""
import os
import sys

import pytest
import torch

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'LoG')), reason='reference tree not present')


@pytest.fixture(scope='module')
def renderer_module():
    cv2 = pytest.importorskip('cv2')  # noqa: F841  (renderer.py imports it)
    sys.path.insert(0, REF)
    try:
        import LoG.render.renderer as R
    finally:
        sys.path.remove(REF)
    return R


def camera():
    return {'FoVx': 1.0, 'FoVy': 0.8, 'image_height': 48, 'image_width': 64, 'world_view_transform': torch.eye(4),
            'full_proj_transform': torch.eye(4), 'camera_center': torch.zeros(3), 'K': torch.eye(3)}


def test_reference_renderer_binds_to_dropin(renderer_module):
    R = renderer_module
    import log_b200.rasterizer as ours
    assert R.GaussianRasterizer is ours.GaussianRasterizer                      # `from diff_gaussian_rasterization_wodilate import`
    R.NaiveRendererAndLoss(use_origin_render=True)
    assert R.BaseRender.GaussianRasterizer is ours.StockGaussianRasterizer      # `from diff_gaussian_rasterization import`
    R.NaiveRendererAndLoss(use_origin_render=False)
    assert R.BaseRender.GaussianRasterizer is ours.GaussianRasterizer
    rast = R.BaseRender.prepare(camera(), torch.zeros(3))
    s = rast.raster_settings
    assert (s.image_width, s.image_height, s.sh_degree, s.prefiltered, s.debug) == (64, 48, 0, False, False)
    assert abs(s.tanfovx - 0.5463024898) < 1e-6


class _Model:
    """The three attributes renderer.render() touches (renderer.py:118-140, 174)."""
    training = False
    visibility_flag = None
    empty_xyz = torch.zeros((0, 3))

    def get_all(self, camera, rasterizer, **kw):
        n = 5
        return {'xyz': torch.rand(n, 3), 'opacity': torch.rand(n, 1), 'colors': torch.rand(n, 3),
                'scaling': torch.rand(n, 3), 'rotation': torch.nn.functional.normalize(torch.rand(n, 4))}


def test_reference_render_reaches_our_forward_with_its_own_kwargs(renderer_module, built):
    """CPU tensors: our forward must be reached (no TypeError on the keyword set, including use_filter=False for the
    fork in eval mode, renderer.py:151-152) and must refuse loudly instead of computing on the CPU."""
    R = renderer_module
    from log_b200._capi import LgrError
    rr = R.NaiveRendererAndLoss(use_origin_render=False)
    rast = R.BaseRender.prepare(camera(), torch.zeros(3))
    with pytest.raises(LgrError, match='no CPU fallback'):
        rr.render(camera(), rast, _Model())
