"""CPU: a derandomised sweep of small random configurations (image size, Gaussian count, splat size, flavour, filter, camera)
through the public API on the SIMT emulation against the C oracle -- binning and sub-tile edge cases (splats on tile borders,
images that are not a multiple of 16 or even smaller than a tile, one-Gaussian scenes) that the hand-picked cases may miss."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, assume, given, settings, strategies as st

import test_gpu_parity as gp
from oracle import torch_dense as O
from util import f32_camera, run_gpu


@settings(derandomize=True, max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
@given(W=st.integers(5, 130), H=st.integers(5, 90), n=st.integers(1, 350), sigma=st.sampled_from([0.4, 1.0, 2.5, 6.0, 15.0]),
       flavour=st.sampled_from(['fork', 'stock']), use_filter=st.booleans(), rot=st.booleans(), seed=st.integers(0, 10_000),
       deg=st.integers(0, 3))
def test_random_small_configurations(emulated_backend, W, H, n, sigma, flavour, use_filter, rot, seed, deg):
    if flavour == 'stock':
        use_filter = True
    # Far-sub-pixel splats WITHOUT the low-pass filter are left out: the backward forms central moments from tensor-core
    # moments about the tile centre, which amplifies the TF32 split's ~2^-22 round-off by (distance to the tile centre / sigma)^2
    # per Gaussian -- <= 5e-5 with the filter on (sigma >= 0.55 px: every training configuration), unbounded as sigma -> 0
    # without it (this sweep found 5e-4 on dscales for a lone splat of sigma = 0.07 px).  LoG switches the filter off only for
    # evaluation (renderer.py:151-152), where no backward runs.  DESIGN.md, parity status.
    assume(use_filter or sigma >= 2.5)
    kwc = dict(R=[[0.98, 0.0, 0.199], [0, 1, 0], [-0.199, 0, 0.98]], T=[0.1, -0.05, 0.3]) if rot else {}
    cam = f32_camera(O.make_camera(W, H, bg=(0.3, 0.1, 0.6), sh_degree=deg, **kwc))
    sc = gp.f32_scene(O.make_scene(n, W, H, sigma, seed=seed, sh_degree=deg))
    G = O.make_cotangent(3, H, W, seed=seed + 1).to(torch.float32).to(torch.float64)
    fm = O.FILTER_ADD if flavour == 'stock' else (O.FILTER_MAX if use_filter else O.FILTER_NONE)
    ref = gp.oracle(cam, sc, G, fm, deg)
    ref32 = gp.oracle(cam, sc, G, fm, deg, dtype=np.float32)
    got = run_gpu(cam, sc, G, flavour=flavour, use_filter=use_filter, sh_degree=deg)
    gs, ps = gp.borderline_decisions(cam, sc, fm, thr=1e-5)          # pairs within fp32 round-off of the alpha = 1/255 rule
    gp.sc_n[0] = n
    gp.check_all(gp.drop(got, gs, ps), gp.drop(ref, gs, ps), deg, flavour == 'fork', H * W, gp.drop(ref32, gs, ps))


@settings(derandomize=True, max_examples=8, deadline=None, suppress_health_check=list(HealthCheck))
@given(world=st.integers(1, 7), W=st.integers(8, 100), H=st.integers(8, 120), n=st.integers(1, 500))
def test_random_shard_configurations(emulated_backend, world, W, H, n):
    """Shard mode (SplatExchange over `world` virtual ranks, two steps through the same buffers) against the single-GPU path:
    image and radii bit-identical, aux outputs equal, gradients within 2e-5 -- for band counts that leave ranks without tile
    rows (world > tile rows), shards without Gaussians (n < world) and regions with no row at all."""
    import shard_checks
    shard_checks.run_two_steps(world, 0, 'fork', size=(W, H, n))


@settings(derandomize=True, max_examples=6, deadline=None, suppress_health_check=list(HealthCheck))
@given(world=st.integers(1, 6), W=st.integers(16, 120), H=st.integers(16, 120), n=st.integers(1, 600))
def test_random_band_mode_configurations(emulated_backend, world, W, H, n):
    """Band mode (round-1 multi-GPU layout): packed gradient rows of `world` tile-row bands reproduce the dense gradients."""
    gp.test_band_mode_rows_reproduce_dense_gradients(True, world, size=(W, H, n))


@settings(derandomize=True, max_examples=6, deadline=None, suppress_health_check=list(HealthCheck))
@given(deg=st.integers(0, 3), W=st.integers(5, 120), H=st.integers(5, 100), N=st.integers(2, 800), frac=st.floats(0.01, 1.0))
def test_random_gather_fused_configurations(emulated_backend, deg, W, H, N, frac):
    """render_gathered (index + raw parameter tables -> projection) against LoG's get_all -> activations -> call sequence."""
    gp.test_gather_fused_render_equals_log_get_all(True, deg, size=(W, H, N, max(1, int(N * frac))))


@settings(derandomize=True, max_examples=8, deadline=None, suppress_health_check=list(HealthCheck))
@given(W=st.integers(5, 120), H=st.integers(5, 100), n=st.integers(0, 600), r=st.sampled_from([0.5, 2.0, 6.0, 20.0]))
def test_random_point_id_count_configurations(emulated_backend, W, H, n, r):
    """point_id / point_count from the blend's winner histogram equal torch.unique over the H x W winner image."""
    gp.test_point_id_count_equals_torch_unique(True, W, H, n, r)
