"""`-m gpu`: the rows built after the round-1 GPU budget was spent -- the fused tree walk and LoG's
colour activation with SH fused into the projection (row 3) -- on hardware, with the same checks that pass on the CPU
emulation (tests/test_tree_traverse.py, tests/test_emulated_host.py).  Green on the B200 since round 2."""
import pytest

import test_gpu_parity as gp
import test_tree_traverse as tt

pytestmark = pytest.mark.gpu


def test_tree_walk_reproduces_the_reference_lists(built):
    tt.check_goldens()


def test_tree_walk_matches_oracle_with_culled_points(built):
    tt.check_culled_scene_against_oracle()


@pytest.mark.parametrize('deg', [1, 3])
def test_fused_log_colour_activation_with_sh(built, deg):
    gp.check_fused_log_colour_activation_with_sh(deg)
