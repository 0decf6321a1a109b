#!/usr/bin/env python
"""Randomised simulation of collect_quad_kernel's mbarrier protocol (CPU only): producer, two MMA issuers and the
two sets of eight epilogue warps as state machines over phase-bit barriers, stepped in random order.  Checks that
the parities and arrival counts used in the kernel neither deadlock nor let an agent touch a ring slot / TMEM
buffer that another one still owns.  Mirrors the indices of the kernel one to one (t & 15, t & 3, (t >> 2) & 1 ...).
usage: quad_protocol_sim.py [n_tiles] [seeds]"""
import random, sys

RING, NBUF, EPI_PER_SET = 16, 4, 8


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def done(self, parity):          # mbarrier.try_wait.parity: has the phase with this parity completed?
        return self.phase != parity


def run(n_visit, seed):
    rnd = random.Random(seed)
    full = [Bar(1) for _ in range(RING)]
    empty = [Bar(1) for _ in range(RING)]
    tfull = [Bar(1) for _ in range(NBUF)]
    tempty = [Bar(EPI_PER_SET) for _ in range(NBUF)]
    slot_tile = [None] * RING        # which tile a ring slot holds
    buf_tile = [None] * NBUF         # which tile's scores a TMEM buffer holds
    buf_reads = [0] * NBUF
    seen = {(s, w): [] for s in (0, 1) for w in range(EPI_PER_SET)}

    def producer():
        for t in range(n_visit):
            s = t & (RING - 1)
            while not empty[s].done(((t // RING) & 1) ^ 1):
                yield
            assert slot_tile[s] is None, f"slot {s} overwritten while tile {slot_tile[s]} is unread"
            slot_tile[s] = t
            full[s].arrive()         # expect_tx + complete_tx of the bulk copy
            yield

    def issuer(i):
        for t in range(i, n_visit, 2):
            s, buf = t & (RING - 1), t & 3
            while not full[s].done((t // RING) & 1):
                yield
            while not tempty[buf].done(((t >> 2) & 1) ^ 1):
                yield
            assert slot_tile[s] == t, f"issuer {i} read slot {s}: holds {slot_tile[s]}, wants {t}"
            assert buf_tile[buf] is None, f"buffer {buf} overwritten while tile {buf_tile[buf]} is unread"
            buf_tile[buf], buf_reads[buf] = t, 0
            yield                    # the MMA runs
            tfull[buf].arrive()      # commit -> accumulators ready
            slot_tile[s] = None
            empty[s].arrive()        # commit -> stage free
            yield

    def epilogue(set_, w):
        for t in range(set_, n_visit, 2):
            buf = t & 3
            while not tfull[buf].done((t >> 2) & 1):
                yield
            assert buf_tile[buf] == t, f"set {set_} warp {w}: buffer {buf} holds {buf_tile[buf]}, wants {t}"
            seen[(set_, w)].append(t)
            yield                    # tcgen05.ld + wait
            buf_reads[buf] += 1
            if buf_reads[buf] == EPI_PER_SET:
                buf_tile[buf] = None
            tempty[buf].arrive()
            for _ in range(rnd.randrange(4)):
                yield                # reduce

    agents = [producer(), issuer(0), issuer(1)] + [epilogue(s, w) for s in (0, 1) for w in range(EPI_PER_SET)]
    live = list(agents)
    idle_rounds = 0
    snapshot = None
    while live:
        a = rnd.choice(live)
        try:
            next(a)
        except StopIteration:
            live.remove(a)
        state = (tuple(b.phase for b in full + empty + tfull + tempty), tuple(b.pending for b in tempty), len(live))
        idle_rounds = idle_rounds + 1 if state == snapshot else 0
        snapshot = state
        assert idle_rounds < 200000, "no progress: deadlock"
    for (s, w), ts in seen.items():
        assert ts == list(range(s, n_visit, 2)), f"set {s} warp {w} saw {ts[:8]}..."


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    for n_visit in (2, 4, 6, 34, n):
        for seed in range(seeds):
            run(n_visit, seed)
    print(f"ok: tiles 2, 4, 6, 34, {n} x {seeds} random schedules, no deadlock, no slot or buffer overwritten early")
