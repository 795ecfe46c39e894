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


class _FakeSlot:
    """Stands in for runner._Slot: records launches, returns the batch index as the 'result'."""
    log = []

    def __init__(self, runner, n, h, w):
        self.shape = (n, h, w)
        self.busy = False
        self.value = None

    def launch(self, runner, left, right):
        assert not self.busy, "a slot was relaunched before its result was taken"
        self.busy = True
        self.value = int(left[0, 0, 0, 0])
        _FakeSlot.log.append(("launch", self.shape, self.value))

    def wait(self):
        import torch
        assert self.busy
        self.busy = False
        _FakeSlot.log.append(("wait", self.shape, self.value))
        return torch.full((self.shape[0], 1, 1, 1), float(self.value))


def test_infer_batches_order_and_slot_reuse(monkeypatch):
    """Results come back in submission order, at most n_streams batches per shape are in flight, and a slot is never
    relaunched before its previous result was handed out — also when input shapes interleave."""
    import numpy as np
    from codd_b200 import runner as R
    monkeypatch.setattr(R, "_Slot", _FakeSlot)
    _FakeSlot.log = []

    class _Stereo:
        def eval(self):
            return None

    r = R.StereoSequenceRunner(_Stereo(), device="cpu", use_graph=False, n_streams=2)
    shapes = [(2, 8, 8), (2, 8, 8), (1, 4, 4), (2, 8, 8), (2, 8, 8), (1, 4, 4), (1, 4, 4), (1, 4, 4), (2, 8, 8)]
    batches = []
    for i, (n, h, w) in enumerate(shapes):
        a = np.zeros((n, h, w, 3), np.uint8)
        a[0, 0, 0, 0] = i
        batches.append((a, a))
    outs = list(r.infer_batches(batches))
    assert [int(o[0, 0, 0, 0]) for o in outs] == list(range(len(shapes)))
    assert [o.shape[0] for o in outs] == [s[0] for s in shapes]
    waits = [e[2] for e in _FakeSlot.log if e[0] == "wait"]
    assert waits == sorted(waits)                       # results are taken in submission order
