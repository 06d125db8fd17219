"""Shard mode (log_b200/sharded.py:SplatExchange, csrc/lgr_shard.cu) emulated on ONE GPU: R virtual ranks inside this
process, the "peer" exchange buffers are plain local tensors and the barriers are no-ops because the phases run one
after the other on one stream.  Every rank owns a block of Gaussians and a band of tile rows; the union of what the
ranks produce must equal the ordinary single-GPU call:

  * image, radii, point_id_pixel, point_weight_pixel, point_weight: EXACTLY (each pixel is blended by one rank from the
    same records in the same (depth, global index) order);
  * gradients: to fp32 summation order (the 2D gradients are accumulated per band and then summed).

Green on the B200 since the first hardware run of round 2 (the xfail marker of round 1 is gone).  The second step of
run_two_steps goes through the same buffers with different data; with `sync_free` switched on it would be a device-sized
step (covered separately by shard_checks.run_two_steps(..., sync_free=True)).
"""
import pytest

import shard_checks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('world,deg,flavour', [(2, 0, 'fork'), (3, 0, 'fork'), (2, 3, 'stock'), (8, 0, 'fork')])
def test_shard_mode_emulated_on_one_gpu(built, world, deg, flavour):
    shard_checks.run_two_steps(world, deg, flavour)


def test_shard_mode_device_sized_steps(built):
    shard_checks.run_device_sized_steps(3)


def test_shard_mode_handles_empty_shards_and_bands(built):
    shard_checks.run_empty_shards_and_bands()
