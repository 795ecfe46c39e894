"""Model check of the rolling-ring convolution's synchronisation protocol (csrc/conv_tc_ring.cu) on the CPU.

The kernel's warp roles talk through mbarriers that are waited on by PARITY (one bit per barrier).  That is only
sound if no waiter can fall two phases behind a barrier — a property of the ring depths and of which warp set
handles which staged row, not of any single line of code.  This test replays the protocol (same barriers, counts,
parities and hand-over counter as the kernel) under many random schedules, with asynchronous TMA completions and an
in-order tensor pipe, and checks the data hazards the barriers exist to prevent:

  * a split group reads fp32 stage g % NBUF only after the TMA load of row g has landed, the TMA overwrites it only
    after the split of row g - NBUF is done;
  * the issuer reads operand stage g % NH only after the split of row g wrote it; the split overwrites it only after
    both passes of row g - NH have EXECUTED;
  * the epilogue drains output row r only after the MMAs of staged rows r, r+1, r+2 executed; pass A overwrites a TMEM
    slot only after the epilogue drained the row that used it RING rows earlier;
  * MMAs enter the tensor queue in row order (bit-deterministic accumulation);  nothing deadlocks.

It also shows why the ring depths must be even (the kernel's static_assert): with an odd fp32 depth the two split
groups see every other phase of a stage barrier and a parity wait can pass two phases early.
"""
import random

import pytest


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.done = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.pending = self.count
            self.done += 1

    def passed(self, parity):           # mbarrier.try_wait.parity: true iff the current phase bit != parity
        return (self.done & 1) != parity


def simulate(rows, nbuf, nh, ring, seed, check=True, slow_row=None):
    rnd = random.Random(seed)
    full = [Bar(1) for _ in range(nbuf)]
    empty = [Bar(1) for _ in range(nbuf)]         # one split group (modelled as one arrival) frees an fp32 stage
    lo = [Bar(1) for _ in range(nh)]
    hempty = [Bar(1) for _ in range(nh)]
    accf = [Bar(1) for _ in range(ring)]
    acce = [Bar(1) for _ in range(ring)]
    st = dict(tma_landed=set(), split_done=set(), mma_exec=set(), drained=set(), issued=[], a_issued=0)
    in_flight = []                                 # TMA loads issued, not yet landed
    queue = []                                     # tensor pipe FIFO: ("mma", g) | ("commit", bar)
    viol = []

    def need(cond, msg):
        if not cond:
            viol.append(msg)

    def producer():
        for g in range(rows):
            sb = g % nbuf
            while not empty[sb].passed(((g // nbuf) & 1) ^ 1):
                yield
            need(g < nbuf or (g - nbuf) in st["split_done"], f"TMA row {g} overwrites a stage still being split")
            in_flight.append(g)
            yield

    def split(group):
        for g in range(group, rows, 2):
            sb, hb = g % nbuf, g % nh
            while not hempty[hb].passed(((g // nh) & 1) ^ 1):
                yield
            while not full[sb].passed((g // nbuf) & 1):
                yield
            need(g in st["tma_landed"], f"split reads row {g} before its TMA load landed")
            need(g < nh or (g - nh) in st["mma_exec"], f"split overwrites operand stage of row {g - nh} before its MMAs ran")
            yield
            st["split_done"].add(g)
            lo[hb].arrive()
            empty[sb].arrive()
            yield

    def issuer(group):
        for g in range(group, rows, 2):
            hb = g % nh
            while not lo[hb].passed((g // nh) & 1):
                yield
            need(g in st["split_done"], f"MMA of row {g} issued before its operand stage was written")
            slot = g % ring                      # fresh output row g overwrites its TMEM slot
            while not acce[slot].passed(((g // ring) & 1) ^ 1):
                yield
            need(g < ring or (g - ring) in st["drained"], f"pass A overwrites the slot of undrained row {g - ring}")
            while st["a_issued"] < g:
                yield
            st["issued"].append(g)
            queue.append(("mma", g))             # pass A then pass B of the row
            st["a_issued"] = g + 1
            queue.append(("commit", hempty[hb]))
            if g >= 2:
                queue.append(("commit", accf[(g - 2) % ring]))
            yield
        # the last two output rows are completed by the halo rows in the kernel; model: commit them at the end
        if group == (rows - 1) % 2:
            while st["a_issued"] < rows:
                yield
            for r in (rows - 2, rows - 1):
                if r >= 0:
                    queue.append(("commit", accf[r % ring]))

    def epilogue():
        for r in range(rows):
            slot = r % ring
            while not accf[slot].passed((r // ring) & 1):
                yield
            need(all(x in st["mma_exec"] for x in (r, r + 1, r + 2) if x < rows), f"epilogue drains row {r} early")
            yield
            st["drained"].add(r)
            acce[slot].arrive()
            yield

    def hardware():
        while True:
            acted = False
            cand = [g for g in in_flight if g != slow_row] or ([] if rnd.random() > 0.01 else list(in_flight))
            if cand and rnd.random() < 0.5:
                g = cand[0] if rnd.random() < 0.8 else rnd.choice(cand)                        # mostly in order
                in_flight.remove(g)
                st["tma_landed"].add(g)
                full[g % nbuf].arrive()
                acted = True
            if queue and rnd.random() < 0.5:
                kind, x = queue.pop(0)
                if kind == "mma":
                    st["mma_exec"].add(x)
                else:
                    x.arrive()
                acted = True
            yield acted

    roles = [producer(), split(0), split(1), issuer(0), issuer(1), epilogue()]
    hw = hardware()
    alive = list(roles)
    quiet = 0                                       # scheduler steps since the hardware last did something
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
    except AssertionError as exc:                   # a barrier received more arrivals than it expects
        if check:
            raise
        viol.append(str(exc))
    if check:
        assert not alive, "deadlock: roles still waiting"
        assert st["issued"] == list(range(rows)), "MMAs were not queued in row order"
    return viol


@pytest.mark.parametrize("nbuf,nh,ring", [(6, 10, 16), (4, 6, 16), (4, 4, 8)])     # the three launch configurations
def test_protocol_is_hazard_free(nbuf, nh, ring):
    for seed in range(60):
        assert simulate(40, nbuf, nh, ring, seed) == []


def test_odd_stage_depth_breaks_the_parity_wait():
    """NBUF = 3: rows g and g + 3 of one fp32 stage belong to DIFFERENT split groups, so a group sees every other phase
    of the stage's barrier; when the load of row 3 lands late, group 0 passes the wait for row 6 on the completion of
    row 0 (found on the GPU as a timing-dependent launch failure before the ring depths were required to be even)."""
    found = any(simulate(40, 3, 4, 8, seed, check=False, slow_row=3) for seed in range(200))
    assert found
    # the same adversarial memory system (one load that lands very late) is harmless with even depths
    for seed in range(40):
        assert simulate(40, 4, 4, 8, seed, slow_row=3) == []
