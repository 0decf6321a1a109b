#!/usr/bin/env python
"""Cycle-stepped timing model of one SM running the hi*hi-only collect pass (CPU only, a planning aid).

Three layouts of the same work (128 queries x 256 domains per "stage", scores to be max-reduced):
  lockstep  two 256-column TMEM buffers, all 16 epilogue warps take 64 columns of every stage     (round 1, v6)
  two_sets  two 256-column buffers, a set of 8 warps per buffer takes 128 columns of every other
            stage in two rounds of loads, buffer handed back after the second round                 (round 1, v7)
  quad      four 128-column buffers, two per set; a warp takes 64 columns of an own 128-column tile
            and hands the buffer back right after its loads                            (collect_quad_kernel)

Resources per SM, with the figures measured in scripts/umma_microbench.cu on B200:
  tensor pipe   one MMA at a time: 171 cycles for M128 N256 K16, 118 for M128 N128 K16
  tcgen05.ld    128 B/clk per scheduler (SMSP), 64 B/clk per warp; a warp's x32 load is 4 KB
  ALU           one warp instruction per 2 cycles per SMSP (FMNMX3 is half rate), round-robin over ready warps
  wake-ups      an mbarrier waiter resumes ~WAKE cycles after the arrival that completes the phase
The warp arbiter is modelled twice: round-robin, and strictly highest-warp-id-first (what the in-kernel traces of
round 1 suggested: the lowest warp of a scheduler reaches a new stage ~480 cycles after the highest).
The model has no hits (theta = +inf): compare with the 57.4 ms pure-scan measurement of the two-set kernel
(~540 cycles per 256 columns at 1.9 GHz, 7750 stages, 26 waves).

usage: collect_timing_model.py [stages]
"""
import sys

WAKE = 90            # mbarrier wake-up latency of a spinning waiter (cycles)
COMMIT = 40          # MMA completion -> barrier phase flip
LD_LAT = 60          # fixed latency of a tcgen05.ld on top of its transfer time
INSTR_PER_CHUNK = 17 # max tree of one 32-column chunk
OVERHEAD = 12        # ALU-pipe instructions per own stage outside the trees (addresses, compares, loop)
CHUNK_BYTES = 32 * 32 * 4


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase, self.flip_at = count, count, 0, None

    def arrive(self, now, delay=0):
        self.pending -= 1
        if self.pending == 0:
            self.pending = self.count
            self.flip_at = now + delay

    def tick(self, now):
        if self.flip_at is not None and now >= self.flip_at:
            self.phase ^= 1
            self.flip_at = None
            self.last_flip = now

    def done(self, parity):
        return self.phase != parity


def simulate(layout, n_stages, n_mma=1, prio=False):
    quad = layout == "quad"
    n_buf = 4 if quad else 2
    cols_per_buf = 128 if quad else 256
    t_mma = (118 if quad else 171) * n_mma
    n_tiles = n_stages * (2 if quad else 1)               # a tile fills one buffer
    warps_per_buf = 8 if layout != "lockstep" else 16
    tfull = [Bar(1) for _ in range(n_buf)]
    tempty = [Bar(warps_per_buf) for _ in range(n_buf)]
    tensor_free_at = [0]

    # ---- issuers: thread i owns tiles i, i+2, ... ----
    class Issuer:
        def __init__(self, i):
            self.t, self.state, self.until = i, "wait", 0

        def step(self, now):
            if self.t >= n_tiles:
                return
            buf = self.t % n_buf
            use = self.t // n_buf
            if self.state == "wait":
                if tempty[buf].done((use & 1) ^ 1):
                    woke = getattr(tempty[buf], "last_flip", -10 ** 9) + WAKE
                    if now >= woke:
                        start = max(now, tensor_free_at[0])
                        tensor_free_at[0] = start + t_mma
                        self.state, self.until = "mma", start + t_mma
            elif now >= self.until:                        # the issuing thread is blocked while the MMA runs
                tfull[buf].arrive(now, COMMIT)
                self.t += 2
                self.state = "wait"

    # ---- epilogue warps ----
    class Warp:
        def __init__(self, w):
            self.w, self.smsp = w, w & 3
            grp = w >> 2
            self.set = grp & 1
            if layout == "lockstep":
                self.tiles = iter(range(n_tiles))
                self.rounds = 1                            # one pair of loads per stage (64 columns)
            elif layout == "two_sets":
                self.tiles = iter(range(self.set, n_tiles, 2))
                self.rounds = 2                            # 128 columns in two pairs of loads
            else:
                self.tiles = iter(range(self.set, n_tiles, 2))
                self.rounds = 1
            self.state, self.tile = "next", None
            self.bytes_left = self.alu_left = 0
            self.round = 0
            self.ready_at = 0
            self.done_tiles = 0

        def want_ld(self):
            return self.state == "ld" and self.bytes_left > 0

        def want_alu(self, now):
            return self.state == "alu" and now >= self.ready_at

        def step(self, now):
            if self.state == "next":
                self.tile = next(self.tiles, None)
                if self.tile is None:
                    self.state = "end"
                    return
                self.state, self.round = "wait", 0
            if self.state == "wait":
                buf, use = self.tile % n_buf, self.tile // n_buf
                if tfull[buf].done(use & 1) and now >= getattr(tfull[buf], "last_flip", -10 ** 9) + WAKE:
                    self.state, self.bytes_left = "ld", 2 * CHUNK_BYTES
            elif self.state == "ld" and self.bytes_left <= 0:
                self.round += 1
                if self.round == self.rounds:              # the warp's share of the buffer is in registers
                    tempty[self.tile % n_buf].arrive(now + LD_LAT)
                self.state = "alu"
                self.ready_at = now + LD_LAT
                self.alu_left = 2 * INSTR_PER_CHUNK + OVERHEAD // self.rounds
            elif self.state == "alu" and self.alu_left <= 0:
                if self.round < self.rounds:
                    self.state, self.bytes_left = "ld", 2 * CHUNK_BYTES
                else:
                    self.done_tiles += 1
                    self.state = "next"

    issuers = [Issuer(0), Issuer(1)]
    warps = [Warp(w) for w in range(16)]
    rr = [0, 0, 0, 0]
    now = 0
    alu_busy = 0
    while any(w.state != "end" for w in warps):
        for b in tfull + tempty:
            b.tick(now)
        for i in issuers:
            i.step(now)
        for w in warps:
            w.step(now)
        for s in range(4):
            mine = [w for w in warps if w.smsp == s]
            loading = [w for w in mine if w.want_ld()]
            if loading:                                    # 128 B/clk per SMSP, 64 B/clk per warp
                share = min(64, 128 // len(loading))
                for w in loading:
                    w.bytes_left -= share
            if now % 2 == 0:                               # one half-rate instruction per two cycles
                ready = [w for w in mine if w.want_alu(now)]
                if ready:
                    pick = max(ready, key=lambda x: x.w) if prio else ready[rr[s] % len(ready)]
                    rr[s] += 1
                    pick.alu_left -= 1
                    alu_busy += 2
        now += 1
        if now > 4000 * n_stages:
            raise RuntimeError("model stuck")
    return now / n_stages, alu_busy / (4.0 * now)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    print("cycles per 256 columns (ALU busy), no hits.  Measured on B200: two_sets hi*hi-only ~540; lockstep full split ~800 (config-4 shape)")
    print(f"{'layout':10s} {'MMAs':>4s} {'fair arbiter':>16s} {'high-id-first arbiter':>24s}")
    for n_mma in (1, 3):
        for layout in ("lockstep", "two_sets", "quad"):
            a, b = simulate(layout, n, n_mma, False), simulate(layout, n, n_mma, True)
            print(f"{layout:10s} {n_mma:4d} {a[0]:10.0f} ({a[1]:.2f}) {b[0]:18.0f} ({b[1]:.2f})")
