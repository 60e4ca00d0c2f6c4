"""Model of the level-set propagation of the GPU (adaptive-sph_b200/csrc/level.cu: k_propagate — front pushes with an
unsigned atomicMin on the bit pattern / on an order-preserving key, first push claims, stop at the cutoff) checked
against the reference's Jacobi sweeps over all particles (simulation.rs:729-801) on random particle clouds.

Pure Python, fp32 via numpy; run:  python tools/level_model.py [cases]
tests/test_level_model.py runs a few cases of it.
"""
import sys

import numpy as np

F = np.float32
UNASSIGNED = 0xFFFFFFFF


def level_key(v):
    """level.cu: level_key — larger float <=> smaller unsigned key, either sign."""
    b = int(np.array([v], dtype=F).view(np.uint32)[0])
    return b if b & 0x80000000 else 0x7FFFFFFF - b


def level_of_key(k):
    b = k if k & 0x80000000 else 0x7FFFFFFF - k
    return np.array([b], dtype=np.uint32).view(F)[0]


class Cloud:
    def __init__(self, rng, n, signed):
        side = np.sqrt(n) * 0.8
        self.n = n
        self.pos = (rng.random((n, 2)) * side).astype(F)
        self.pos[:, 1] *= F(0.5)  # a slab: the surface is its top edge, the propagation runs many sweeps deep
        h = (F(0.9) + rng.random(n).astype(F) * F(0.5))
        self.neigh = []
        for i in range(n):
            d = self.pos - self.pos[i]
            d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
            s = (h + h[i]) * F(0.5) * F(1.4)
            self.neigh.append([int(j) for j in np.nonzero(d2 < s * s)[0] if j != i])
        top = self.pos[:, 1].max()
        self.surface = self.pos[:, 1] > top - F(0.6)
        # EmptyAngle: surface value 0; CenterDiff: values of either sign
        self.phi0 = np.where(self.surface, (rng.normal(0, 0.3, n).astype(F) if signed else np.zeros(n, F)), F(0)).astype(F)

    def dist(self, i, j):
        d = self.pos[j] - self.pos[i]
        return np.sqrt(F(d[0] * d[0]) + F(d[1] * d[1]), dtype=F)


def reference_sweeps(c):
    """simulation.rs:739-800: every sweep recomputes every particle that has no value yet from the neighbours that had
    one BEFORE the sweep; until nothing changes.  Returns (has value, value, number of sweeps)."""
    have = c.surface.copy()
    val = c.phi0.copy()
    sweeps = 0
    changed = True
    while changed:
        changed = False
        have2, val2 = have.copy(), val.copy()
        for i in range(c.n):
            if have[i]:
                continue
            best = None
            for j in c.neigh[i]:
                if have[j]:
                    est = F(val[j] - c.dist(j, i))
                    best = est if best is None else max(best, est)
            if best is not None:
                have2[i], val2[i] = True, best
                changed = True
        have, val = have2, val2
        sweeps += 1
    return have, val, sweeps


def front_pushes(c, signed, dmax=None):
    """level.cu k_propagate: sweep t pushes from the particles assigned in sweep t - 1 into neighbours that are unassigned
    or were claimed in this very sweep; the stored word only ever shrinks (atomicMin on the bit pattern, values <= 0, or on
    the key, either sign).  With dmax: stop after a sweep that assigned only values below -dmax."""
    key = (lambda v: level_key(v)) if signed else (lambda v: int(np.array([v], dtype=F).view(np.uint32)[0]))
    word = [key(c.phi0[i]) if c.surface[i] else UNASSIGNED for i in range(c.n)]
    stamp = [0 if c.surface[i] else -1 for i in range(c.n)]
    front = [i for i in range(c.n) if c.surface[i]]
    sweeps, t = 0, 0
    live_prev = True
    while True:
        t += 1
        if not front or (t > 1 and not live_prev):
            break
        sweeps = t
        new, live = [], False
        for j in front:
            lj = level_of_key(word[j]) if signed else np.array([word[j]], dtype=np.uint32).view(F)[0]
            for i in c.neigh[j]:
                if stamp[i] == -1 or stamp[i] == t:
                    v = F(lj - c.dist(j, i))
                    word[i] = min(word[i], key(v))
                    if stamp[i] == -1:
                        stamp[i] = t
                        new.append(i)
                    if dmax is None or v > -dmax:
                        live = True
        front, live_prev = new, live
    have = np.array([w != UNASSIGNED for w in word])
    val = np.array([(level_of_key(w) if signed else np.array([w], dtype=np.uint32).view(F)[0]) if w != UNASSIGNED else F(0) for w in word], dtype=F)
    return have, val, max(1, sweeps)


def clamped(have, val, dmax):
    """what the smoothing pass reads (simulation.rs:833-836)"""
    return np.where(have, np.maximum(val, -dmax), -dmax).astype(F)


def run_case(seed, n=220):
    rng = np.random.default_rng(seed)
    signed = bool(seed & 1)
    c = Cloud(rng, n, signed)
    h0, v0, s0 = reference_sweeps(c)
    h1, v1, s1 = front_pushes(c, signed)
    assert np.array_equal(h0, h1), "assigned sets differ"
    assert np.array_equal(v0.view(np.uint32)[h0], v1.view(np.uint32)[h1]), "values differ"
    assert s0 == s1, (s0, s1)
    if not signed:
        assert (v0[h0] <= 0).all()
        dmax = F(abs(float(v0[h0].min())) * 0.45)  # a cutoff that really cuts
        h2, v2, s2 = front_pushes(c, signed, dmax)
        assert s2 <= s1
        assert np.array_equal(clamped(h0, v0, dmax).view(np.uint32), clamped(h2, v2, dmax).view(np.uint32)), "cutoff changes what the smoothing reads"
        return s1, s2
    return s1, s1


def key_order_ok(rng, count=4000):
    v = np.concatenate([rng.normal(0, 1, count), [0.0, -0.0, 1e-30, -1e-30, 3e38, -3e38]]).astype(F)
    k = np.array([level_key(x) for x in v], dtype=np.uint64)
    o = np.argsort(v, kind="stable")
    ks = k[o]
    vs = v[o]
    for a in range(len(vs) - 1):
        if vs[a] < vs[a + 1] and not ks[a] > ks[a + 1]:
            return False
    return all(level_of_key(int(kk)) == x or (x == 0 and level_of_key(int(kk)) == 0) for kk, x in zip(k, v))


if __name__ == "__main__":
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    for seed in range(cases):
        print(seed, run_case(seed))
    print("keys", key_order_ok(np.random.default_rng(0)))
