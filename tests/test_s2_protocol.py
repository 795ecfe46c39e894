"""Model check of the strided tcgen05 kernel's synchronisation protocol (csrc/conv_tc_s2.cu, fp16-operand version) on
the CPU, in the style of test_ring_protocol.py.

Roles and barriers (all waited on by one-bit parity):
  producer (TMA):        waits EMPTY[s], loads stage s, FULL[s] completes when the bytes land;
  conversion groups 0/1: take ALTERNATE stages; wait FULL[s], rewrite the stage in place as fp16 hi | lo, arrive LO[s];
  MMA issuer:            waits LO[s] (and ACCE[a] before the first stage of a tile), queues pass A and pass B, commits
                         EMPTY[s]; after the 4th stage of a tile commits ACCF[a];
  epilogue:              waits ACCF[a], drains the accumulator, arrives ACCE[a].
Checked: the conversion reads a stage only after its load landed; the TMA overwrites a stage only after the MMAs that read
it EXECUTED; the MMAs of a stage are issued only after its conversion; an accumulator is overwritten only after the
epilogue drained the tile that used it; nothing deadlocks.  With an ODD number of stage buffers the two conversion groups
see every other phase of a stage's FULL barrier and the parity wait can pass a phase early — found on the GPU as a hang
of the bench (5 buffers), now a static_assert in the kernel; the test reproduces it."""
import random

import pytest


class Bar:
    def __init__(self, count=1):
        self.count, self.pending, self.done = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.pending = self.count
            self.done += 1

    def passed(self, parity):           # mbarrier.try_wait.parity: true iff the current phase bit != parity
        return (self.done & 1) != parity


def simulate(tiles, nbuf, nacc, seed, check=True, slow_stage=None):
    rnd = random.Random(seed)
    stages = 4 * tiles                              # one stage per kernel row ky, four per output tile
    full = [Bar() for _ in range(nbuf)]
    lo = [Bar() for _ in range(nbuf)]               # (one arrival per group of 128 threads, modelled as one)
    empty = [Bar() for _ in range(nbuf)]
    accf = [Bar() for _ in range(nacc)]
    acce = [Bar() for _ in range(nacc)]
    st = dict(landed=set(), converted=set(), mma_exec=set(), drained=set())
    in_flight, queue, viol = [], [], []

    def need(cond, msg):
        if not cond:
            viol.append(msg)

    def producer():
        for it in range(stages):
            sb = it % nbuf
            while not empty[sb].passed(((it // nbuf) & 1) ^ 1):
                yield
            need(it < nbuf or (it - nbuf) in st["mma_exec"], f"TMA stage {it} overwrites operands the MMAs still read")
            in_flight.append(it)
            yield

    def convert(group):
        for it in range(stages):
            if (it & 1) != group:
                continue
            sb = it % nbuf
            while not full[sb].passed((it // nbuf) & 1):
                yield
            need(it in st["landed"], f"conversion reads stage {it} before its load landed")
            yield
            st["converted"].add(it)
            lo[sb].arrive()
            yield

    def issuer():
        for it in range(stages):
            sb, ti, ky = it % nbuf, it >> 2, it & 3
            ab = ti % nacc
            while not lo[sb].passed((it // nbuf) & 1):
                yield
            need(it in st["converted"], f"MMAs of stage {it} issued before its conversion")
            if ky == 0:
                while not acce[ab].passed(((ti // nacc) & 1) ^ 1):
                    yield
                need(ti < nacc or (ti - nacc) in st["drained"], f"tile {ti} overwrites the accumulator of undrained tile {ti - nacc}")
            queue.append(("mma", it))
            queue.append(("commit", empty[sb]))
            if ky == 3:
                queue.append(("commit", accf[ab]))
            yield

    def epilogue():
        for ti in range(tiles):
            ab = ti % nacc
            while not accf[ab].passed((ti // nacc) & 1):
                yield
            need(all(4 * ti + k in st["mma_exec"] for k in range(4)), f"epilogue drains tile {ti} early")
            yield
            st["drained"].add(ti)
            acce[ab].arrive()
            yield

    def hardware():
        while True:
            acted = False
            cand = [g for g in in_flight if g != slow_stage] or ([] if rnd.random() > 0.01 else list(in_flight))
            if cand and rnd.random() < 0.5:
                g = cand[0] if rnd.random() < 0.8 else rnd.choice(cand)
                in_flight.remove(g)
                st["landed"].add(g)
                full[g % nbuf].arrive()
                acted = True
            if queue and rnd.random() < 0.5:
                kind, x = queue.pop(0)                # the tensor pipe executes in order
                if kind == "mma":
                    st["mma_exec"].add(x)
                else:
                    x.arrive()
                acted = True
            yield acted

    alive = [producer(), convert(0), convert(1), issuer(), epilogue()]
    hw = hardware()
    quiet = 0
    try:
        while alive and quiet < 50000:
            if rnd.random() < 0.3:
                quiet = 0 if next(hw) else quiet + 1
                continue
            r = rnd.choice(alive)
            try:
                next(r)
            except StopIteration:
                alive.remove(r)
            quiet += 1
        while in_flight or queue:
            next(hw)
    except AssertionError as exc:
        if check:
            raise
        viol.append(str(exc))
    if check:
        assert not alive, "deadlock: roles still waiting"
    elif alive:
        viol.append("deadlock")
    return viol


@pytest.mark.parametrize("nbuf,nacc", [(10, 4), (8, 4), (4, 4), (2, 4), (12, 4)])      # the launch configurations
def test_s2_protocol_is_hazard_free(nbuf, nacc):
    for seed in range(60):
        assert simulate(12, nbuf, nacc, seed) == []


def test_s2_odd_stage_depth_breaks_the_parity_wait():
    """5 stage buffers: stages it and it + 5 of one buffer belong to different conversion groups; when one load lands late
    a group passes its FULL wait on a phase that completed two uses earlier (the hang seen on the GPU)."""
    found = any(simulate(12, 5, 4, seed, check=False, slow_stage=5) for seed in range(300))
    assert found
    for seed in range(40):
        assert simulate(12, 4, 4, seed, slow_stage=5) == []
