"""Executable model of k_bisect's work distribution (arrowhead.jl_b200/csrc/ah_pipe.cuh, comment above k_bisect).

The device code hands eigenpairs to slots by ticket and time-slices them through a FIFO in global memory.  Its
correctness argument -- every eigenpair is bisected to the end exactly once, no slot ever blocks another, the launch
terminates -- does not depend on the arithmetic, so it is checked here on the CPU: the protocol is restated step by
step (one memory operation per micro-step, as the device executes it) and run under random interleavings of the
slots, including the windows in which a ticket number has been taken but the FIFO entry is not written yet.
"""
import random

import pytest

FREE, BISECT, PENDING, DONE = range(4)


class Launch:
    """Global memory of one k_bisect round-0 launch: dcount[1] tickets, dcount[4] requeues, dcount[5] finished, queue."""

    def __init__(self, steps, nslots, quantum, window, qcap, ready=None):
        self.cnt = len(steps)
        self.need = list(steps)                 # bisection steps each eigenpair needs (device: until the bracket converges)
        self.progress = [0] * self.cnt          # KState.steps: written back on a requeue, read by the next slot
        self.ready = ready or [True] * self.cnt  # KState.flag == KS_READY (k_prep hands the others to k_full)
        self.tickets = 0
        self.nreq = 0
        self.done = 0
        self.queue = {}                         # r -> column (entry carries the launch tag: unwritten entries do not match)
        self.nslots, self.quantum, self.window, self.qcap = nslots, quantum, window, qcap
        self.finalized = [0] * self.cnt
        self.holders = [0] * self.cnt           # slots currently working on the eigenpair (must never exceed 1)


class Slot:
    """Warp j of a CTA at a sweep boundary.  `micro()` performs ONE global-memory operation (or one sweep) and returns."""

    def __init__(self, launch):
        self.L = launch
        self.phase = FREE
        self.col = None
        self.local = 0          # S.steps
        self.q = 0              # steps since the slot took this eigenpair
        self.ticket = None
        self.pc = "boundary"
        self.tmp = None

    def micro(self):
        L = self.L
        if self.phase == DONE:
            return
        if self.pc == "boundary":
            if self.phase == BISECT:
                if self.local >= L.need[self.col]:          # bisect_converged -> bisect_finalize
                    L.finalized[self.col] += 1
                    L.holders[self.col] -= 1
                    self.pc = "count_done"
                    return
                left = L.need[self.col] - self.local        # (device: estimated from the bracket width)
                qeff = min(L.quantum, max(1, left >> 2)) if L.quantum > 0 else 0
                if L.quantum > 0 and self.q >= qeff:
                    self.q = 0
                    self.pc = "swap_read_head"
                    return
                self.pc = "sweep"
                return
            if self.phase == FREE:
                self.ticket = L.tickets                      # atomicAdd(tickets, 1)
                L.tickets += 1
                self.phase = PENDING
                return
            if self.phase == PENDING:
                t = self.ticket
                if t < L.cnt:
                    col = t
                elif L.quantum == 0:
                    self.phase = DONE
                    return
                else:
                    r = t - L.cnt
                    if r < L.qcap and r in L.queue:          # entry with this launch's tag
                        col = L.queue[r]
                    else:
                        self.pc = "pending_read_done"        # (a second, later load)
                        return
                if not L.ready[col]:
                    self.pc = "count_done"
                    return
                self.col, self.local, self.q = col, L.progress[col], 0
                L.holders[col] += 1
                assert L.holders[col] == 1, "two slots hold the same eigenpair"
                self.phase = BISECT
                return
        elif self.pc == "pending_read_done":
            if L.done >= L.cnt:                              # everything finished: nothing will be queued any more
                assert self.ticket - L.cnt not in L.queue, "an eigenpair was queued for a slot that retires"
                self.phase, self.pc = DONE, "boundary"
            else:
                self.pc = "idle_sweep"                       # the CTA runs its sweep without this slot
        elif self.pc == "count_done":
            if L.quantum > 0:
                L.done += 1                                  # atomicAdd(dcount + 5, 1)
            self.phase, self.col, self.pc = FREE, None, "boundary"
        elif self.pc == "swap_read_head":
            self.tmp = L.tickets                             # volatile load
            self.pc = "swap_read_nreq"
        elif self.pc == "swap_read_nreq":
            head, nreq = self.tmp, L.nreq                    # (the two loads are not atomic together)
            tail = L.cnt + nreq
            swap = head < tail and tail - head < L.window * L.nslots and nreq + L.nslots < L.qcap
            self.pc = "swap_writeback" if swap else "sweep"
        elif self.pc == "swap_writeback":
            L.progress[self.col] = self.local                # G.left / right / steps, then __threadfence
            self.pc = "swap_take_r"
        elif self.pc == "swap_take_r":
            self.tmp = L.nreq                                # atomicAdd(dcount + 4, 1)
            L.nreq += 1
            assert self.tmp < L.qcap, "requeue buffer overrun"
            self.pc = "swap_publish"
        elif self.pc == "swap_publish":
            L.holders[self.col] -= 1
            L.queue[self.tmp] = self.col                     # the entry becomes visible only now
            self.phase, self.col, self.pc = FREE, None, "boundary"
        elif self.pc == "sweep":
            self.local += 1
            self.q += 1
            self.pc = "boundary"
        elif self.pc == "idle_sweep":
            self.pc = "boundary"


def run(steps, nslots, quantum, window, qcap, seed, ready=None, max_micro=5_000_000):
    rng = random.Random(seed)
    L = Launch(steps, nslots, quantum, window, qcap, ready)
    slots = [Slot(L) for _ in range(nslots)]
    live = list(range(nslots))
    n = 0
    while live:
        n += 1
        assert n < max_micro, "no termination"
        # adversarial-ish scheduling: mostly uniform, sometimes one slot runs a burst while the others stand still
        j = rng.choice(live)
        for _ in range(rng.choice((1, 1, 1, 2, 5, 40))):
            slots[j].micro()
        if slots[j].phase == DONE:
            live.remove(j)
    return L


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("cfg", [
    dict(cnt=40, nslots=16, quantum=8, window=2),        # 2.5 eigenpairs per slot: time slicing from the first sweep on
    dict(cnt=200, nslots=16, quantum=8, window=2),       # long launch: first-come-first-served, slicing in the last waves
    dict(cnt=37, nslots=16, quantum=1, window=1000),     # pathological: requeue after every step
    dict(cnt=10, nslots=16, quantum=8, window=2),        # fewer eigenpairs than slots
    dict(cnt=64, nslots=8, quantum=0, window=2),         # no time slicing (round 1, small blocks)
])
def test_every_eigenpair_is_finished_exactly_once(cfg, seed):
    rng = random.Random(1000 + seed)
    steps = [rng.randint(45, 62) for _ in range(cfg["cnt"])]
    ready = [rng.random() > 0.05 for _ in range(cfg["cnt"])]          # a few are handed to k_full by k_prep
    L = run(steps, cfg["nslots"], cfg["quantum"], cfg["window"], qcap=cfg["cnt"] * 24 + 64, seed=seed, ready=ready)
    for c in range(cfg["cnt"]):
        assert L.finalized[c] == (1 if ready[c] else 0)
        assert L.holders[c] == 0
    if cfg["quantum"] > 0:
        assert L.done == cfg["cnt"]
    assert L.tickets >= cfg["cnt"] + L.nreq                           # every queued entry was claimed by a ticket


@pytest.mark.parametrize("seed", range(6))
def test_full_requeue_buffer_stops_time_slicing_not_the_launch(seed):
    """When the FIFO is (nearly) full the slots simply keep their eigenpairs: nothing is lost, nothing overruns."""
    rng = random.Random(seed)
    cnt, nslots = 60, 8
    steps = [rng.randint(45, 62) for _ in range(cnt)]
    L = run(steps, nslots, quantum=1, window=1000, qcap=nslots + 25, seed=seed)
    assert all(f == 1 for f in L.finalized)
    assert L.nreq < nslots + 25


def test_time_slicing_levels_the_finishing_times():
    """What it is for: with 1.7 eigenpairs per slot first-come-first-served needs two full waves of sweeps, time slicing
    about sum(steps) / slots (counted in sweeps of the slowest slot; uniform scheduling = all slots at the same pace)."""
    def sweeps(quantum):
        steps = [54] * 27
        L = Launch(steps, 16, quantum, 2, 27 * 24 + 64)
        slots = [Slot(L) for _ in range(16)]
        rounds = 0
        while any(s.phase != DONE for s in slots):
            rounds += 1
            for s in slots:                                           # one sweep boundary + one sweep per slot and round
                for _ in range(40):
                    was_sweep = s.pc in ("sweep", "idle_sweep")
                    s.micro()
                    if was_sweep or s.phase == DONE:
                        break
            assert rounds < 10_000
        return rounds
    fcfs, sliced = sweeps(0), sweeps(8)
    assert fcfs >= 2 * 54
    assert sliced <= 27 * 54 / 16 + 12
