"""Model of the event-driven greedy partner search (adaptive-sph_b200/csrc/adapt.cu: k_greedy) checked against the
reference's serial loop (particle_sharing.rs:34-104 / particle_merging.rs:43-115) on random particle clouds.

Pure Python, fp32 via numpy scalars; run:  python tools/greedy_model.py [cases]
The CUDA kernel follows this file phase by phase; tests/test_greedy_model.py runs a few cases of it.
"""
import sys

import numpy as np

F = np.float32
AVAILABLE, DELETE = 0xFFFFFFFF, 0xFFFFFFFE
TOO_SMALL, SMALL, OPTIMAL, LARGE, TOO_LARGE = 0, 1, 2, 3, 4
PENDING, DONE = 1, 2
NONE = -1


class Cloud:
    def __init__(self, rng, n, merging, flags):
        self.n, self.merging = n, merging
        self.pos = rng.random((n, 2)).astype(F) * F(np.sqrt(n) * 0.9)
        self.h = (F(1.0) + rng.random(n).astype(F) * F(0.6))
        self.mass = (F(0.3) + rng.random(n).astype(F) * F(1.2))
        self.target = (F(0.8) + rng.random(n).astype(F) * F(1.0))
        self.mass_base = F(2.2)
        self.dist_factor = F(1.0 + rng.random())
        self.dt = F(0.3)
        self.rate = F(2.0)
        self.cls = np.empty(n, dtype=np.int64)
        for i in range(n):
            r = self.mass[i] / self.target[i]
            self.cls[i] = TOO_SMALL if r <= F(0.5) else SMALL if r <= F(1.0) / F(1.1) else OPTIMAL if r < F(1.1) else LARGE if r < F(2.0) else TOO_LARGE
        self.flags = flags  # allow_merge_optimal, allow_share_optimal, allow_share_too_small, allow_merge_size_diff
        # symmetric neighbour lists incl. self, ascending index (the oracle's order)
        self.neigh = []
        for i in range(n):
            d = self.pos - self.pos[i]
            d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
            s = (self.h + self.h[i]) * F(0.5) * F(2.0)
            self.neigh.append([int(j) for j in np.nonzero(d2 < s * s)[0]])
        self.donor_class = TOO_SMALL if merging else LARGE

    def dropped(self, i):
        if self.merging:
            return self.mass[i]
        return min(self.mass[i] - self.target[i], self.target[i] * self.rate * self.dt)

    def static_eligible(self, d, j):
        am_o, as_o, as_ts, am_sd = self.flags
        cj = self.cls[j]
        if self.merging:
            can = False if cj in (LARGE, TOO_LARGE) else (am_o if cj == OPTIMAL else True)
            if am_sd and self.mass[j] > F(5.0) * self.mass[d]:
                can = True
        else:
            can = True if cj == SMALL else as_ts if cj == TOO_SMALL else as_o if cj == OPTIMAL else False
        if not can:
            return False
        dx = self.pos[d] - self.pos[j]
        md = (self.h[d] + self.h[j]) * F(0.5) * self.dist_factor
        return not (dx[0] * dx[0] + dx[1] * dx[1] > md * md)

    def mass_ok(self, j, add):
        nm = self.mass[j] + add
        return not (nm >= self.target[j] * F(1.1)) and not (nm > self.mass_base)


def serial(c):
    partner = [AVAILABLE] * c.n
    counter = [0] * c.n
    for i in range(c.n):
        if c.cls[i] != c.donor_class:
            continue
        for j in c.neigh[i]:
            if j == i or not c.static_eligible(i, j):
                continue
            if not c.mass_ok(j, c.dropped(i) / F(counter[i] + 1)):
                continue
            if partner[j] != AVAILABLE:
                continue
            if counter[i] == 0:
                if partner[i] != AVAILABLE:
                    continue
                partner[i] = DELETE
            partner[j] = i
            counter[i] += 1
    return partner, counter


def inner_loop(c, d, partner):
    """The donor's loop over its neighbours in ascending index; partner is None for the optimistic run (every
    receiver taken as available).  Returns the claimed receivers."""
    claimed = []
    for j in c.neigh[d]:
        if j == d or not c.static_eligible(d, j):
            continue
        if not c.mass_ok(j, c.dropped(d) / F(len(claimed) + 1)):
            continue
        if partner is not None and partner[j] != AVAILABLE:
            continue
        claimed.append(j)
    return claimed


def event_driven(c, rng):
    n = c.n
    partner = [AVAILABLE] * n
    counter = [0] * n
    state = [0] * n
    nopt = [0] * n
    head = [NONE] * n
    nxt = [NONE] * n
    resume = [0] * n
    work = []
    for d in range(n):  # phase I
        if c.cls[d] != c.donor_class:
            continue
        nopt[d] = len(inner_loop(c, d, None))
        if nopt[d] == 0:
            state[d] = DONE
        else:
            state[d] = PENDING
            work.append(d)

    def touches(y, x):  # x in C^(y): y itself, or a receiver y may still claim
        if x == y:
            return True
        return c.static_eligible(y, x) and c.mass_ok(x, c.dropped(y) / F(nopt[y]))

    def members(d):  # enumeration of C^(d): index 0 = d, 1 + k = list position k
        yield 0, d
        for k, x in enumerate(c.neigh[d]):
            if x != d and touches(d, x):
                yield 1 + k, x

    rounds = 0
    while work:
        rounds += 1
        rng.shuffle(work)
        ready = []
        for d in work:  # phase A: frozen state
            if partner[d] != AVAILABLE:
                state[d] = DONE
                continue
            blocker = NONE
            for q, x in members(d):
                if q < resume[d]:
                    continue
                for y in c.neigh[x]:
                    if c.cls[y] == c.donor_class and y < d and state[y] == PENDING and partner[y] == AVAILABLE and touches(y, x):
                        blocker = y
                        break
                if blocker != NONE:
                    resume[d] = q
                    break
            if blocker == NONE:
                ready.append(d)
            else:
                nxt[d] = head[blocker]
                head[blocker] = d
        work = []
        rng.shuffle(ready)
        for d in ready:  # phase B: ready donors touch disjoint particle sets
            claimed = inner_loop(c, d, partner)
            if claimed:
                assert partner[d] == AVAILABLE
                partner[d] = DELETE
            for j in claimed:
                assert partner[j] == AVAILABLE
                partner[j] = d
            counter[d] = len(claimed)
            state[d] = DONE
            for b in [d] + claimed:
                z = head[b]
                head[b] = NONE
                while z != NONE:
                    work.append(z)
                    z2 = nxt[z]
                    nxt[z] = NONE
                    z = z2
    return partner, counter, rounds


def run_case(seed, n=300):
    rng = np.random.default_rng(seed)
    merging = bool(seed & 1)
    flags = tuple(bool(rng.integers(2)) for _ in range(4))
    c = Cloud(rng, n, merging, flags)
    p0, c0 = serial(c)
    p1, c1, rounds = event_driven(c, rng)
    assert p0 == p1, (seed, "partner")
    assert c0 == c1, (seed, "counter")
    return sum(c0), rounds, sum(1 for i in range(n) if c.cls[i] == c.donor_class)


if __name__ == "__main__":
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    for s in range(cases):
        claims, rounds, donors = run_case(s)
        print(f"seed {s}: {'merge' if s & 1 else 'share'} donors {donors} claims {claims} rounds {rounds}")
    print("ok")
