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


def test_sequence_stats_mean_std_and_dump(tmp_path):
    """Per-sequence rows aggregated like the reference's RunningStats (utils/running_stats.py): mean and (population) std
    over sequences; checked against the reference class itself when /root/reference (or its mirror) is importable."""
    import numpy as np
    from codd_b200.metrics import SequenceStats
    st = SequenceStats()
    rows = [dict(epe=1.0, th3=0.5), dict(epe=2.0, th3=0.25), dict(epe=4.5, th3=0.0)]
    for i, r in enumerate(rows):
        st.push(f"s{i}", r)
    assert st.n == 3
    assert abs(st.mean["epe"] - 2.5) < 1e-12 and abs(st.std["epe"] - np.std([1.0, 2.0, 4.5])) < 1e-12
    st.dump(str(tmp_path / "stats.csv"))
    lines = open(tmp_path / "stats.csv").read().strip().splitlines()
    assert lines[0] == "name,epe,th3" and lines[-2].startswith("mean,2.5") and len(lines) == 6
    try:
        from oracle import ref_loader
        if not ref_loader.available():
            return
        ref_loader.load()
        from utils.running_stats import RunningStats
    except Exception:
        return
    rs = RunningStats()
    for r in rows:
        rs.push(np.array([r["epe"], r["th3"]]))
    assert np.allclose(rs.mean, [st.mean["epe"], st.mean["th3"]], atol=1e-6)
    assert np.allclose(rs.std, [st.std["epe"], st.std["th3"]], atol=1e-6)
