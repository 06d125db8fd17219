"""CPU: the host-side mirror of the reference interface (names, arguments, error behaviour); no kernels run."""
import inspect

import pytest
import torch


def test_settings_fields_match_reference_kwargs():
    """kwargs at LoG/render/renderer.py:63-76."""
    from log_b200 import GaussianRasterizationSettings
    assert list(GaussianRasterizationSettings._fields) == [
        'image_height', 'image_width', 'tanfovx', 'tanfovy', 'bg', 'scale_modifier', 'viewmatrix', 'projmatrix',
        'sh_degree', 'campos', 'prefiltered', 'debug']


def test_rasterizer_call_signature_matches_reference_call_sites():
    """renderer.py:141-153 passes these keywords; level_of_gaussian.py:59 calls .compute_radius."""
    from log_b200 import GaussianRasterizer
    params = inspect.signature(GaussianRasterizer.forward).parameters
    for k in ('means3D', 'means2D', 'shs', 'colors_precomp', 'opacities', 'scales', 'rotations', 'cov3D_precomp', 'use_filter'):
        assert k in params
    assert hasattr(GaussianRasterizer, 'compute_radius')
    s = _settings()
    r = GaussianRasterizer(raster_settings=s)
    for k in ('projmatrix', 'viewmatrix', 'tanfovx', 'tanfovy', 'image_width', 'image_height'):   # level_of_gaussian.py:73-78
        assert hasattr(r.raster_settings, k)


def _settings():
    from log_b200 import GaussianRasterizationSettings
    return GaussianRasterizationSettings(image_height=32, image_width=32, tanfovx=0.5, tanfovy=0.5, bg=torch.zeros(3),
                                         scale_modifier=1.0, viewmatrix=torch.eye(4), projmatrix=torch.eye(4), sh_degree=0,
                                         campos=torch.zeros(3), prefiltered=False, debug=False)


def test_dropin_module_names_import():
    import diff_gaussian_rasterization as stock
    import diff_gaussian_rasterization_wodilate as fork
    assert stock.GaussianRasterizer.flavour == 'stock' and fork.GaussianRasterizer.flavour == 'fork'
    assert stock.GaussianRasterizationSettings is fork.GaussianRasterizationSettings


def test_cpu_tensors_fail_loudly(built):
    """No CPU fallback: CPU tensors must raise, not silently compute."""
    from log_b200 import GaussianRasterizer, compute_radius
    from log_b200._capi import LgrError
    r = GaussianRasterizer(_settings())
    n = 4
    with pytest.raises(LgrError):
        r(means3D=torch.zeros(n, 3), means2D=torch.zeros(n, 3), shs=None, colors_precomp=torch.zeros(n, 3),
          opacities=torch.zeros(n, 1), scales=torch.ones(n, 3), rotations=torch.ones(n, 4), cov3D_precomp=None)
    with pytest.raises(LgrError):
        compute_radius(torch.zeros(n, 3), torch.ones(n, 3), torch.ones(n, 4), torch.eye(4), torch.eye(4), 1., 1., 1., 1.)


def test_argument_validation_mirrors_reference():
    from log_b200 import GaussianRasterizer
    r = GaussianRasterizer(_settings())
    z = torch.zeros(2, 3)
    with pytest.raises(Exception):      # both colour sources (stock raises the same way)
        r(means3D=z, means2D=z, shs=torch.zeros(2, 1, 3), colors_precomp=z, opacities=z[:, :1], scales=z, rotations=torch.zeros(2, 4))
    with pytest.raises(Exception):      # neither
        r(means3D=z, means2D=z, shs=None, colors_precomp=None, opacities=z[:, :1], scales=z, rotations=torch.zeros(2, 4))
    # scales/rotations XOR cov3D_precomp, the stock rule (diff_gaussian_rasterization's forward raises the same sentence)
    with pytest.raises(Exception, match='exactly one of either scale/rotation pair or precomputed 3D covariance'):
        r(means3D=z, means2D=z, shs=None, colors_precomp=z, opacities=z[:, :1], scales=z, rotations=torch.zeros(2, 4), cov3D_precomp=torch.zeros(2, 6))
    with pytest.raises(Exception, match='exactly one of either scale/rotation pair or precomputed 3D covariance'):
        r(means3D=z, means2D=z, shs=None, colors_precomp=z, opacities=z[:, :1], scales=None, rotations=None, cov3D_precomp=None)
    with pytest.raises(NotImplementedError):      # LoG's fused activations act on scales / rotations
        r(means3D=z, means2D=z, shs=None, colors_precomp=z, opacities=z[:, :1], scales=None, rotations=None, cov3D_precomp=torch.zeros(2, 6),
          raw_params=True)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from log_b200 import _capi
    monkeypatch.setattr(_capi, '_lib', None)
    monkeypatch.setattr(_capi, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_capi.LgrError):
        _capi.load()
