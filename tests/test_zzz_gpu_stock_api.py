"""`-m gpu`: members of the stock `diff_gaussian_rasterization` class that LoG itself never uses (cov3D_precomp, markVisible),
added after the round-2 GPU budget was spent: green on the CPU emulation (tests/test_emulated_host.py, also under its
AddressSanitizer build), first hardware run = the round-end suite.  The file sorts last so that nothing else waits on it."""
import pytest

import test_gpu_parity as gp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('flavour', ['stock', 'fork'])
def test_cov3D_precomp_matches_oracle_and_the_scale_rotation_path(built, flavour):
    gp.check_cov3D_precomp(flavour=flavour)


def test_mark_visible(built):
    gp.check_mark_visible()
