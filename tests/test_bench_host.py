"""CPU: host-side arithmetic of bench.py -- the algorithmic-byte model is SURVEY.md section 8(d)'s, to the digit."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


@pytest.mark.parametrize('n,H,W,deg,b_min_gb', [
    (10_000_000, 1080, 1920, 0, 1.946),       # 10 M, 1080p, precomputed colour
    (100_000, 1080, 1920, 3, 0.139),          # 100 k, 1080p, SH-3
    (10_000_000, 1080, 1920, 3, 7.346),       # 10 M, 1080p, SH-3
    (50_000_000, 2160, 3840, 0, 9.665),       # 50 M, 4K
])
def test_algorithmic_bytes_reproduce_the_survey_table(n, H, W, deg, b_min_gb):
    import bench
    per_kernel, b_min, b_model = bench.algorithmic_bytes(n, H, W, 0, deg)
    assert abs(b_min / 1e9 - b_min_gb) < 5e-4
    assert b_model == b_min                                     # no instances: B_model = B_min
    D = 20_594_520                                              # + 112 B per (Gaussian, tile) instance: 32 sort + 40 + 40 blend
    per_kernel_d, b_min_d, b_model_d = bench.algorithmic_bytes(n, H, W, D, deg)
    assert b_min_d == b_min and b_model_d - b_min == 112 * D
    assert sum(per_kernel_d.values()) == b_model_d
    assert per_kernel_d['bin_sort'] == 32 * D
    assert per_kernel_d['blend_fwd'] - per_kernel['blend_fwd'] == 40 * D == per_kernel_d['blend_bwd'] - per_kernel['blend_bwd']


def test_workloads_are_the_baseline_configs():
    import bench
    assert bench.WORKLOADS['10m'][:3] == (10_000_000, 1920, 1080) and bench.WORKLOADS['10m'][4] == 0
    assert bench.WORKLOADS['100k'][:3] == (100_000, 1920, 1080) and bench.WORKLOADS['100k'][4] == 3
    assert bench.WORKLOADS['1k'][:3] == (1_000, 256, 256)
    assert bench.WORKLOADS['50m4k'][:3] == (50_000_000, 3840, 2160)
