"""N2 host logic (CPU): sequence sharding follows DistributedSampler(shuffle=False) of apis/inference.py."""
import pytest

from codd_b200.runner import shard_indices


def test_shard_indices_partition():
    for n in (0, 1, 5, 8, 13):
        for world in (1, 2, 3, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert all(p == list(range(r, n, world)) for r, p in enumerate(parts))
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)
