"""`-m gpu`: the fused tree walk on hardware against the reference's own lists and the oracle (same checks as the CPU
emulation runs in tests/test_tree_traverse.py).  Written after the round-1 GPU budget was spent: non-strict xfail until
its first hardware run (expected: XPASS); sorts last so that it cannot disturb the verified tests."""
import pytest

import test_tree_traverse as tt

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason='tree walk: first hardware run pending')]


def test_tree_walk_reproduces_the_reference_lists(built):
    tt.check_goldens()


def test_tree_walk_matches_oracle_with_culled_points(built):
    tt.check_culled_scene_against_oracle()
