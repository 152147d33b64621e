"""Model check of the producer / consumer hand-over of the warp-specialised callback kernel (`ws_mix_tile`,
oddio_b200/csrc/odb_scene_mix.cu), on the CPU: the same index arithmetic (team -> source sequence, slot ring, wrap
counts, mbarrier phase parities, entries per producer pass), agents stepped by a random scheduler.

Checked for random (sources, batch size, grid, ring depth, pass size, tiles): nobody deadlocks; every consumer warp
sees exactly the shipped kernel's source sequence for its team (batches gp + r GP, sources in order), each source once;
a slot is never rewritten before both consumer warps of the team have handed it back; a parity wait never passes on a
phase other than the one it means. This pins the protocol's arithmetic - a ring of PASS + 1 slots is the minimum (the
kernel's static_assert asks for PASS + 2), a ring of PASS slots deadlocks and the model says so - independently of
the GPU run (`tests/test_kernel_shapes_gpu.py` is the test of the kernel itself)."""
import random

import pytest


class Bar:
    """mbarrier: `count` arrivals complete a phase; try_wait.parity(p) passes once the phase of parity p is over,
    i.e. when the current phase's parity differs from p."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passed(self, parity):
        return (self.phase & 1) != parity


class Cta:
    def __init__(self, block, grid, n_sources, bsz, cwarps, S, PASS):
        self.block, self.G, self.n, self.bsz, self.S, self.PASS = block, grid, n_sources, bsz, S, PASS
        self.teams = cwarps // 2
        self.GP = grid * self.teams
        self.full = [[Bar(1) for _ in range(S)] for _ in range(self.teams)]
        self.empty = [[Bar(2) for _ in range(S)] for _ in range(self.teams)]
        self.slot_src = [[None] * S for _ in range(self.teams)]       # what the producer put there
        self.slot_use = [[-1] * S for _ in range(self.teams)]         # which use of the slot that was
        self.holders = [[0] * S for _ in range(self.teams)]           # consumer warps that have not handed it back
        self.seen = {}                                                # (team, part) -> sources consumed

    def src(self, team, r, q):
        return (self.block * self.teams + team + r * self.GP) * self.bsz + q

    def expected(self, team):
        """The shipped kernel's order: batches gp, gp + GP, ...; inside a batch the sources in order."""
        out, gp = [], self.block * self.teams + team
        n_batches = (self.n + self.bsz - 1) // self.bsz
        b = gp
        while b < n_batches:
            out += [s for s in range(b * self.bsz, b * self.bsz + self.bsz) if s < self.n]
            b += self.GP
        return out


def consumer(cta, team, part, st):
    """One tile of a consumer warp (generator: yields while it waits). st = [slot, parity], carried over tiles."""
    S, seen = cta.S, cta.seen.setdefault((team, part), [])
    slot, par = st
    r = q = 0
    if cta.src(team, r, q) < cta.n:
        use = [0]  # phase index this wait means, for the ABA check
        while not cta.full[team][slot].passed(par):
            yield
        while True:
            q1, r1 = q + 1, r
            if q1 == cta.bsz:
                q1, r1 = 0, r + 1
            more = cta.src(team, r1, q1) < cta.n
            nslot, npar = slot + 1, par
            if nslot == S:
                nslot, npar = 0, par ^ 1
            if more:
                while not cta.full[team][nslot].passed(npar):
                    yield
            # consume: the slot must hold this entry, produced for exactly this use of the slot
            assert cta.slot_src[team][slot] == cta.src(team, r, q), "slot holds another source"
            assert cta.full[team][slot].phase == cta.slot_use[team][slot] + 1, "parity wait passed on the wrong phase"
            seen.append(cta.slot_src[team][slot])
            yield
            cta.holders[team][slot] -= 1
            cta.empty[team][slot].arrive()
            slot, par = nslot, npar
            if not more:
                break
            q, r = q1, r1
            del use
            use = None
    st[0], st[1] = slot, par


def producer(cta, pw, st):
    """One tile of a producer warp. st = [slotA, useA, slotB, useB], carried over tiles."""
    S, E, bsz = cta.S, cta.PASS, cta.bsz
    teams = (2 * pw, 2 * pw + 1)
    slot, use = [st[0], st[2]], [st[1], st[3]]

    def src_of(t, r, q, j):
        qq, rr = q + j, r
        while qq >= bsz:
            qq, rr = qq - bsz, rr + 1
        return cta.src(teams[t], rr, qq)

    r = q = 0
    while True:
        n = [sum(1 for j in range(E) if src_of(t, r, q, j) < cta.n) for t in (0, 1)]
        assert n[1] <= n[0]
        if n[0] == 0:
            break
        entries = []
        for t in (0, 1):
            for j in range(n[t]):
                sl, us = slot[t] + j, use[t]
                if sl >= S:
                    sl, us = sl - S, us + 1
                entries.append((t, j, sl, us))
        for t, j, sl, us in entries:  # both consumers have handed the slot back
            if us > 0:
                while not cta.empty[teams[t]][sl].passed((us - 1) & 1):
                    yield
        for t, j, sl, us in entries:
            assert cta.holders[teams[t]][sl] == 0, "slot rewritten while a consumer still holds it"
            cta.slot_src[teams[t]][sl] = src_of(t, r, q, j)
            cta.slot_use[teams[t]][sl] = us
            cta.holders[teams[t]][sl] = 2
        yield  # (the chain walk)
        for t, j, sl, us in entries:
            cta.full[teams[t]][sl].arrive()
        for t in (0, 1):
            slot[t] += n[t]
            if slot[t] >= S:
                slot[t], use[t] = slot[t] - S, use[t] + 1
        q += E
        while q >= bsz:
            q, r = q - bsz, r + 1
    st[0], st[1], st[2], st[3] = slot[0], use[0], slot[1], use[1]


def run_cta(rng, n_sources, bsz, grid, block, cwarps, S, PASS, tiles):
    cta = Cta(block, grid, n_sources, bsz, cwarps, S, PASS)
    cstate = {(t, p): [0, 0] for t in range(cta.teams) for p in (0, 1)}
    pstate = {pw: [0, 0, 0, 0] for pw in range(cwarps // 4)}
    for _ in range(tiles):
        cta.seen.clear()
        agents = [consumer(cta, t, p, cstate[(t, p)]) for t in range(cta.teams) for p in (0, 1)]
        agents += [producer(cta, pw, pstate[pw]) for pw in range(cwarps // 4)]
        idle = 0
        while agents:
            a = rng.choice(agents)
            before = [(b.phase, b.pending) for row in cta.full + cta.empty for b in row]
            try:
                next(a)
            except StopIteration:
                agents.remove(a)
                idle = 0
                continue
            after = [(b.phase, b.pending) for row in cta.full + cta.empty for b in row]
            idle = 0 if after != before else idle + 1
            assert idle < 200 * (len(agents) + 1), "deadlock: every agent only waits"
        for t in range(cta.teams):  # (the tile barrier of the kernel: everybody is done before the next tile starts)
            for p in (0, 1):
                assert cta.seen.get((t, p), []) == cta.expected(t)


@pytest.mark.parametrize("seed", range(6))
def test_handover_protocol(seed):
    rng = random.Random(seed)
    for _ in range(25):
        cwarps = rng.choice([8, 12, 16])
        PASS = rng.choice([1, 2, 4])
        S = PASS + 2 + rng.randrange(0, 4)
        bsz = rng.choice([1, 3, 5, 7, 8])
        grid = rng.choice([1, 2, 5])
        n_sources = rng.choice([0, 1, 7, 8, 9, 63, 200, 777, 1500])
        run_cta(rng, n_sources, bsz, grid, rng.randrange(grid), cwarps, S, PASS, tiles=rng.choice([1, 2, 3]))


def test_a_ring_of_pass_slots_deadlocks():
    """The model is sensitive: the consumer holds its slot while it waits for the next one, so S = PASS cannot work."""
    for S in (1, 2, 4):
        with pytest.raises(AssertionError, match="deadlock"):
            run_cta(random.Random(1), 777, 8, 1, 0, 16, S, S, tiles=1)
        run_cta(random.Random(1), 777, 8, 1, 0, 16, S + 1, S, tiles=1)


def test_shipped_experiment_shapes():
    rng = random.Random(99)
    for cwarps, S in ((16, 7), (12, 8)):  # SmxWs, SmxWs12 (PASS = 4)
        for n_sources in (12000 // 94, 1184, 4000):
            run_cta(rng, n_sources, 8, 2, 1, cwarps, S, 4, tiles=2)
