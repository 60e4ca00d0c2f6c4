"""One whole physics step of the oracle against an independent numpy restatement of the step semantics (SURVEY.md
Appendix A, written from the formulas, not from oracle/): h, N_2 by brute force, boundary lambda terms from the closed
forms, CFL dt, density, a_ii, non-pressure acceleration, PPE sources, the relaxed-Jacobi update, pressure acceleration and
the integrators of HybridDFSPH / IISPH / OnlyDivergence.  `max_iters: 0` makes every solve exactly one sweep (p = 0 -> a^p = 0
-> p' = max(0, omega * s / a_ii)), so the whole step is a closed expression.  fp64 oracle, tolerance 1e-9 of the field scale:
what differs is only the order of the sums.  (The reference itself cannot be run here)"""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ETA = 1.9


def _w(r, h):
    q = r / (2 * h)
    w = np.where(q < 0.5, 6 * (q ** 3 - q ** 2) + 1, np.where(q < 1, 2 * (1 - q) ** 3, 0.0))
    return 10.0 / (7 * np.pi * h * h) * w


def _gradw(dx, h):
    """cubic_kernel_2d_deriv (A1): zero for q <= 1e-5"""
    r = np.linalg.norm(dx, axis=-1)
    q = r / (2 * h)
    dw = np.where(q < 0.5, 18 * q * q - 12 * q, np.where(q < 1, -6 * (1 - q) ** 2, 0.0))
    with np.errstate(invalid="ignore", divide="ignore"):
        g = (10.0 / (7 * np.pi * h * h) * dw / (2 * h))[..., None] * dx / r[..., None]
    return np.where((q > 1e-5)[..., None], g, 0.0)


class Restatement:
    def __init__(self, lib, params, pos, vel, mass, probes):
        v = params.values
        self.rho0, self.nu, self.g, self.omega = float(v["rest_density"]), float(v["viscosity"]), float(v["gravity"]), float(v["jacobi_omega"])
        self.cfl, self.max_dt, self.hyb = float(v["cfl_factor"]), float(v["max_dt"]), float(v["hybrid_dfsph_factor"])
        self.eps = float(v["sdf_gradient_eps"])
        self.x, self.v, self.m = pos.astype(np.float64), vel.astype(np.float64), mass.astype(np.float64)
        n = len(self.m)
        self.h = ETA * np.sqrt(self.m / self.rho0 / np.pi)
        d = self.x[:, None, :] - self.x[None, :, :]
        hij = 0.5 * (self.h[:, None] + self.h[None, :])
        self.nb = (d ** 2).sum(-1) < (2 * hij) ** 2            # A3, f = 2, strict, self included
        self.W = np.where(self.nb, _w(np.sqrt((d ** 2).sum(-1)), hij), 0.0)
        self.G = np.where(self.nb[..., None], _gradw(d, hij), 0.0)   # gradW_ij, [n, n, 2]
        self.d = d
        self.hij = hij
        # boundary terms (A6), Quadratic1 penalty, lambda from the closed forms the Maxima tables pin
        lib.asph_lambda.restype = C.c_double; lib.asph_lambda.argtypes = [C.c_double]
        lib.asph_dlambda.restype = C.c_double; lib.asph_dlambda.argtypes = [C.c_double]
        self.Lam = np.zeros(n); self.GLam = np.zeros((n, 2))
        for probe in probes:   # signed distance functions, positive inside the tank (sdf/sdf_plane.rs, sdf/sdf2d.rs)
            sr = 2 * self.h
            for i in range(n):
                di = probe(self.x[i]) / sr[i]
                if not di < 1:
                    continue
                ex, ey = np.array([self.eps, 0.0]), np.array([0.0, self.eps])   # central differences, sdf/sdf.rs:50-62
                g = np.array([probe(self.x[i] + ex) - probe(self.x[i] - ex), probe(self.x[i] + ey) - probe(self.x[i] - ey)]) / (2 * self.eps)
                if np.linalg.norm(g) < 1e-5:
                    continue
                g = g / np.linalg.norm(g)
                pen, pder = (1.0, 0.0) if di > 0 else ((0.5 * di * di + 1, di) if di > -1 else (0.5 - di, -1.0))
                lam, lamd = (1.0, 0.0) if di <= -1 else (lib.asph_lambda(di), lib.asph_dlambda(di))
                self.Lam[i] += lam * pen
                self.GLam[i] += g / sr[i] * (pder * lam + pen * lamd)

    def dt(self):
        c = (2 * self.h) ** 2 / ((self.v ** 2).sum(1) + 0.01)
        return min(self.max_dt, self.cfl * np.sqrt(c.min()))

    def density(self):
        return (self.m[None, :] * self.W).sum(1) + self.Lam

    def aii(self, rho):
        S = (self.m[None, :, None] * self.G).sum(1)
        Q = (self.m[None, :] * (self.G ** 2).sum(-1)).sum(1)
        B = self.rho0 * self.GLam / (rho ** 2)[:, None]
        return ((S / (rho ** 2)[:, None] + B) * (S / rho[:, None] + self.rho0 * self.GLam / rho[:, None])).sum(1) + self.m * Q / rho ** 3

    def non_pressure(self, v, rho):
        vij = v[:, None, :] - v[None, :, :]
        xv = (self.d * vij).sum(-1)
        rho_ij = 0.5 * (rho[:, None] + rho[None, :])
        coef = 8.0 * (self.m[None, :] / rho_ij) * xv / ((self.d ** 2).sum(-1) + 0.01 * self.hij ** 2)
        coef = np.where(self.nb & (xv < 0), coef, 0.0)
        a = self.nu * (coef[..., None] * self.G).sum(1)
        a[:, 1] += self.g
        return a

    def divergence(self, q, rho):
        dq = q[None, :, :] - q[:, None, :]
        s = ((self.m[None, :] / rho[:, None]) * (dq * self.G).sum(-1)).sum(1)
        return s + (self.rho0 / rho) * ((0.0 - q) * self.GLam).sum(1)

    def pressure_accel(self, p, rho):
        P = p / rho ** 2
        a = -((self.m[None, :] * (P[:, None] + P[None, :]))[..., None] * self.G).sum(1)
        return a - (self.rho0 * P)[:, None] * self.GLam

    def one_sweep(self, s, aii):
        with np.errstate(divide="ignore", invalid="ignore"):
            p = np.where(np.abs(aii) < 10e-4, 0.0, self.omega * s / aii)
        return np.maximum(p, 0.0)

    def solve(self, s, aii, rho, dt, tol, density_mode, max_iters):
        """iisph_pressure_iterations (A13, sim.rs:1378-1516): relaxed Jacobi from p = 0; returns (p, iterations k)."""
        p = np.zeros(len(s))
        k = 0
        while True:
            ap = self.pressure_accel(p, rho)
            Ap = self.divergence(ap, rho)
            with np.errstate(divide="ignore", invalid="ignore"):
                pn = p + self.omega * (s - Ap) / aii
            singular = np.abs(aii) < 10e-4
            perr = (rho * dt * dt if density_mode else dt) * (s - Ap)
            normal = ~singular & (pn > 0)
            p = np.where(normal, pn, 0.0)
            avg = perr[normal].sum() / normal.sum() if normal.any() else np.nan
            if not normal.any():
                break
            if k > 1 and (abs(avg / self.rho0) < tol if density_mode else abs(avg) < tol / dt):
                break
            if k == max_iters:
                break
            k += 1
        return p, k

    def step_converged(self, params):
        """HybridDFSPH with the full solver loops (default tolerances)."""
        v0 = params.values
        dt = self.dt()
        rho = self.density()
        aii = self.aii(rho)
        v = self.v + dt * self.non_pressure(self.v, rho)
        p, k_div = self.solve(-self.divergence(v, rho) / dt, aii, rho, dt, float(v0["hybrid_dfsph_max_avg_divergence_error"]), False, int(v0["max_iters"]))
        v = v + dt * self.pressure_accel(p, rho)
        s = -(self.rho0 - rho) / (rho * dt * dt) - self.divergence(v, rho) / dt
        p, k_den = self.solve(s, aii, rho, dt, float(v0["hybrid_dfsph_max_avg_density_error"]), True, int(v0["max_iters"]))
        ap = self.pressure_accel(p, rho)
        return {"position": self.x + dt * v + dt * dt * ap, "velocity": v + dt * ap * min(dt * self.hyb, 1.0), "pressure": p,
                "div_iterations": k_div, "density_iterations": k_den}

    def step(self, solver):
        dt = self.dt()
        rho = self.density()
        aii = self.aii(rho)
        v = self.v + dt * self.non_pressure(self.v, rho)
        out = {"dt": dt, "density": rho, "aii": aii}
        if solver == "HybridDFSPH":
            p = self.one_sweep(-self.divergence(v, rho) / dt, aii)
            v = v + dt * self.pressure_accel(p, rho)
            s = -(self.rho0 - rho) / (rho * dt * dt) - self.divergence(v, rho) / dt
            p = self.one_sweep(s, aii)
            ap = self.pressure_accel(p, rho)
            x = self.x + dt * v + dt * dt * ap
            v = v + dt * ap * min(dt * self.hyb, 1.0)
        else:
            s = -self.divergence(v, rho) / dt
            if solver == "IISPH":
                s = s - (self.rho0 - rho) / (rho * dt * dt)
            p = self.one_sweep(s, aii)
            ap = self.pressure_accel(p, rho)
            v = v + dt * ap
            x = self.x + dt * v
        out.update(ppe_source_term=s, pressure=p, pressure_accel=ap, position=x, velocity=v)
        return out


def _polygon_probe(pts):
    """Sdf2D of a closed polygon (sdf/sdf2d.rs:73-141): distance to the nearest edge or vertex, positive on the left of the
    edges (inside a counter-clockwise tank); at a vertex the sign comes from the sum of the two edge normals."""
    pts = np.asarray(pts, np.float64)
    nxt = np.roll(pts, -1, axis=0)
    edir = (nxt - pts) / np.linalg.norm(nxt - pts, axis=1)[:, None]
    elen2 = ((nxt - pts) ** 2).sum(1)
    prev = np.roll(edir, 1, axis=0)
    pn = np.stack([-prev[:, 1] - edir[:, 1], prev[:, 0] + edir[:, 0]], axis=1)

    def probe(x):
        best, out = np.inf, 0.0
        for k in range(len(pts)):
            pd = x - pts[k]
            proj = pd @ edir[k]
            dl = pd[0] * -edir[k, 1] + pd[1] * edir[k, 0]
            if proj > 0 and proj * proj < elen2[k] and dl * dl < best:
                best, out = dl * dl, dl
            c = pd @ pd
            if c < best:
                best, out = c, np.sqrt(c) * (1.0 if pd @ pn[k] >= 0 else -1.0)
        return out
    return probe


@pytest.mark.parametrize("solver", ["HybridDFSPH", "IISPH", "OnlyDivergence"])
@pytest.mark.parametrize("where", ["corner", "middle", "polygon-corner"])
def test_one_step_equals_the_numpy_restatement(asph, oracle64, default_params, solver, where):
    rng = np.random.default_rng(7)
    sp = 0.05
    corner = where != "middle"  # in the corner two walls contribute lambda terms; in the middle none does
    kind = "AnalyticUnderestimate" if where == "polygon-corner" else "AnalyticOverestimate"
    sc = asph.SceneConfig.dam_break(sp, pos=(-0.999, -0.999) if corner else (-0.3, -0.3), size=(0.6, 0.5))
    pos, vel, mass = asph.scene_particles(sc)
    pos = (pos + rng.uniform(-0.15, 0.15, pos.shape) * sp).astype(np.float32)
    if corner:
        pos = np.maximum(pos, np.float32(-0.999))
    mass = (mass * np.exp(rng.uniform(-0.4, 0.4, mass.shape))).astype(np.float32)   # mixed smoothing lengths
    vel = (rng.standard_normal(vel.shape) * 0.2).astype(np.float32)
    params = default_params.replace(merging=False, sharing=False, splitting=False, level_estimation_method="None", max_iters=0,
                                    pressure_solver_method=solver, hybrid_dfsph_factor=30.0, init_boundary_handler=kind)
    b = asph.scene_boundary(sc, kind)
    if kind == "AnalyticOverestimate":
        probes = [(lambda x, k=k: float(b.planes[k][0]) * x[0] + float(b.planes[k][1]) * x[1] + float(b.planes[k][2])) for k in range(b.n_planes)]
    else:
        probes = [_polygon_probe([(float(b.poly[k][0]), float(b.poly[k][1])) for k in range(b.n_poly)])]
    sim = asph.FluidSimulation(params, pos, vel, mass, b, lib=oracle64)
    dt = sim.single_step_without_adaptivity()
    ref = Restatement(oracle64, params, pos, vel, mass, probes).step(solver)
    assert abs(dt - ref["dt"]) <= 1e-7 * ref["dt"]            # the ABI reports dt as a float
    if corner:
        assert np.abs(sim.get_field("lambda_sum")).max() > 0.05   # the walls are really felt
    info = sim.step_info()
    assert info["density_sweeps"] in (0, 1) and info["div_sweeps"] in (0, 1)
    for name in ("density", "aii", "ppe_source_term", "pressure", "pressure_accel", "position", "velocity"):
        got = sim.get_field(name).astype(np.float64)
        scale = max(np.abs(ref[name]).max(), 1e-12)
        # fields come back through the float ABI: 1e-7 relative; the lambda LUT adds ~1e-8 near the walls
        assert np.abs(got - ref[name]).max() <= 3e-7 * scale, (name, np.abs(got - ref[name]).max() / scale)
    sim.close()


def test_solver_loops_equal_the_numpy_restatement(asph, oracle64, default_params):
    """The relaxed-Jacobi loops with their stop rules (A13): same iteration counts, same pressures, same end state."""
    rng = np.random.default_rng(11)
    sp = 0.05
    sc = asph.SceneConfig.dam_break(sp, pos=(-0.999, -0.999), size=(0.6, 0.5))
    pos, vel, mass = asph.scene_particles(sc)
    pos = np.maximum((pos + rng.uniform(-0.15, 0.15, pos.shape) * sp).astype(np.float32), np.float32(-0.999))
    vel = (rng.standard_normal(vel.shape) * 0.2).astype(np.float32)
    params = default_params.replace(merging=False, sharing=False, splitting=False, level_estimation_method="None")
    b = asph.scene_boundary(sc, "AnalyticOverestimate")
    probes = [(lambda x, k=k: float(b.planes[k][0]) * x[0] + float(b.planes[k][1]) * x[1] + float(b.planes[k][2])) for k in range(b.n_planes)]
    sim = asph.FluidSimulation(params, pos, vel, mass, b, lib=oracle64)
    sim.single_step_without_adaptivity()
    ref = Restatement(oracle64, params, pos, vel, mass, probes).step_converged(params)
    info = sim.step_info()
    assert (info["div_iterations"], info["density_iterations"]) == (ref["div_iterations"], ref["density_iterations"]), (info, ref)
    assert info["density_iterations"] > 3   # the loop really iterated
    for name in ("pressure", "position", "velocity"):
        got = sim.get_field(name).astype(np.float64)
        assert np.abs(got - ref[name]).max() <= 1e-6 * max(np.abs(ref[name]).max(), 1e-12), name
    sim.close()


def test_level_set_equals_the_numpy_restatement(asph, oracle64, default_params):
    """EmptyAngle surface detection (A4), propagation by Jacobi sweeps (A5) and the smoothing pass (sim.rs:804-857) on the
    two-resolution scene, against numpy: surface flags, sweep count and the final level field."""
    sc = asph.SceneConfig.from_yaml(os.path.join(ROOT, "configs", "default-scene.yaml"))
    pos, vel, mass = asph.scene_particles(sc)
    rng = np.random.default_rng(3)
    pos = (pos + rng.uniform(-0.1, 0.1, pos.shape) * 0.03).astype(np.float32)   # break the lattice ties of the cone test
    params = default_params.replace(merging=False, sharing=False, splitting=False)
    b = asph.scene_boundary(sc, "AnalyticOverestimate")
    sim = asph.FluidSimulation(params, pos, vel, mass, b, lib=oracle64)
    sim.single_step_without_adaptivity()
    x, m = pos.astype(np.float64), mass.astype(np.float64)
    rho0, D = 1.0, float(params["maximum_surface_distance"])
    h = ETA * np.sqrt(m / rho0 / np.pi)
    d = x[:, None, :] - x[None, :, :]
    r = np.sqrt((d ** 2).sum(-1))
    hij = 0.5 * (h[:, None] + h[None, :])
    f1 = float(np.float64(params["level_estimation_range"])) / ETA
    N1 = r ** 2 < (hij * f1) ** 2
    G = np.where(N1[..., None], _gradw(d, hij), 0.0)
    normal = -(m / rho0)[:, None] * G.sum(1)
    planes = [tuple(float(b.planes[k][c]) for c in range(3)) for k in range(b.n_planes)]
    wall = np.min([nx * x[:, 0] + ny * x[:, 1] + dl for nx, ny, dl in planes], axis=0)
    cos50 = np.cos(np.deg2rad(50.0))
    n = len(m)
    surface = np.zeros(n, bool)
    margin = np.full(n, np.inf)   # distance of the deciding comparison from its threshold (fp32 inputs: skip exact ties)
    for i in range(n):
        js = np.nonzero(N1[i])[0]
        nn = (normal[i] ** 2).sum()
        if len(js) < 3:
            surface[i] = True
        elif nn < 1e-5:
            surface[i] = False; margin[i] = abs(nn - 1e-5)
        elif wall[i] < 1.5 * h[i]:          # boundary_is_fluid_surface: false
            surface[i] = False; margin[i] = abs(wall[i] - 1.5 * h[i])
        else:
            nh = normal[i] / np.sqrt(nn)
            xji = x[js] - x[i]
            c = (xji / (np.linalg.norm(xji, axis=1) + 1e-6)[:, None]) @ nh
            surface[i] = not np.any(c > cos50)
            margin[i] = np.abs(c - cos50).min()
    flags = sim.get_field("flag_is_fluid_surface").astype(bool)
    clear = margin > 1e-6
    assert clear.mean() > 0.99 and np.array_equal(flags[clear], surface[clear])
    surface = flags.copy()   # continue from the oracle's flags so that a tie cannot shift the distance field
    # A5: Jacobi sweeps until nothing changes
    has = surface.copy(); phi = np.zeros(n)
    sweeps = 0
    while True:
        sweeps += 1
        new_has, new_phi, changed = has.copy(), phi.copy(), False
        for i in np.nonzero(~has)[0]:
            js = np.nonzero(N1[i] & has)[0]
            if len(js):
                new_phi[i] = np.max(phi[js] - r[i, js]); new_has[i] = True; changed = True
        has, phi = new_has, new_phi
        if not changed:
            break
    assert sim.step_info()["level_sweeps"] == sweeps
    # smoothing: post-advection positions, N_2 of the step, the step's densities
    x2 = sim.get_field("position").astype(np.float64); rho = sim.get_field("density").astype(np.float64)
    N2 = r ** 2 < (2 * hij) ** 2
    d2 = x2[:, None, :] - x2[None, :, :]
    W = np.where(N2, _w(np.sqrt((d2 ** 2).sum(-1)), hij), 0.0)
    dist = np.where(has, np.maximum(phi, -D), -D)
    vw = (m / rho)[None, :] * W
    level = (vw * dist[None, :]).sum(1) / vw.sum(1)
    got = sim.get_field("level").astype(np.float64)
    assert np.abs(got - level).max() <= 1e-6 * max(np.abs(level).max(), 1e-9), np.abs(got - level).max()
    sim.close()


AVAILABLE, DELETE = 0xFFFFFFFF, 0xFFFFFFFE


def _resample(params, split_patterns, x, v, m, level, h, rows, dt, step_number):
    """single_step_adaptivity restated from SURVEY.md A16-A20 in plain Python (serial greedy searches, swap-with-last
    deletion, append-at-end splitting).  Arrays are float64 copies; returns the new (x, v, m) and (shared, merged, split)."""
    P = params.values
    rho0, D = float(P["rest_density"]), float(P["maximum_surface_distance"])
    r_f, r_b = float(P["particle_radius_fine"]), float(P["particle_radius_base"])
    m_base = np.pi * r_b * r_b * rho0
    x, v, m, level = x.copy(), v.copy(), m.copy(), level.copy()

    def target(i):
        t = max(level[i], -D) / -D
        tr = r_f * (1 - t) + r_b * t
        return np.pi * tr * tr * rho0          # sizing_function: Radius

    def classes():
        out = []
        for i in range(len(m)):
            q = m[i] / target(i)
            out.append(0 if q <= 0.5 else 1 if q <= 1 / 1.1 else 2 if q < 1.1 else 3 if q < 2 else 4)
        return out

    def drop_share(i):
        tm = target(i)
        return min(m[i] - tm, tm * float(P["max_mass_transfer_sharing"]) * dt)

    def search(merging, cls):
        n = len(m)
        partner, counter, claims = [AVAILABLE] * n, [0] * n, 0
        factor = float(P["max_merge_distance"] if merging else P["max_share_distance"])
        for i in range(n):
            counter[i] = 0
            if cls[i] != (0 if merging else 3):
                continue
            for j in rows[i]:
                j = int(j)
                if j == i:
                    continue
                if merging:
                    ok = {3: False, 4: False, 2: bool(P["allow_merge_with_optimal_particle"])}.get(cls[j], True)
                    if P["allow_merge_on_size_difference"] and m[j] > 5 * m[i]:
                        ok = True
                else:
                    ok = {1: True, 0: bool(P["allow_share_with_too_small_particle"]), 2: bool(P["allow_share_with_optimal_particle"])}.get(cls[j], False)
                if not ok:
                    continue
                dx = x[i] - x[j]
                md = 0.5 * (h[i] + h[j]) * factor
                if dx @ dx > md * md:
                    continue
                dropped = m[i] if merging else drop_share(i)
                nm = m[j] + dropped / (counter[i] + 1)
                if nm >= target(j) * 1.1 or nm > m_base or partner[j] != AVAILABLE:
                    continue
                if counter[i] == 0:
                    if partner[i] != AVAILABLE:
                        continue
                    partner[i] = DELETE
                partner[j] = i
                counter[i] += 1
                claims += 1
        return partner, counter, claims

    def receive(merging, partner, counter, minp):
        donors_drop = {d: (m[d] if merging else drop_share(d)) for d in range(len(m)) if partner[d] == DELETE}
        m_old, x_old, v_old = m.copy(), x.copy(), v.copy()
        for i in range(len(m)):
            d = partner[i]
            if d in (AVAILABLE, DELETE) or counter[d] < minp:
                continue
            mu = donors_drop[d] / counter[d]
            tot = m_old[i] + mu
            v[i] = (m_old[i] * v_old[i] + mu * v_old[d]) / tot
            x[i] = (m_old[i] * x_old[i] + mu * x_old[d]) / tot
            m[i] = tot
        return donors_drop

    stats = [0, 0, 0]
    if P["sharing"]:
        partner, counter, stats[0] = search(False, classes())
        drops = receive(False, partner, counter, int(P["minimum_share_partners"]))
        for d, dr in drops.items():
            if counter[d] >= int(P["minimum_share_partners"]):
                m[d] -= dr
    if step_number % 2 == 0:
        if P["merging"]:
            partner, counter, stats[1] = search(True, classes())
            receive(True, partner, counter, int(P["minimum_merge_partners"]))
            order = list(range(len(m)))      # swap-with-last compaction on an index list
            i, last = 0, len(order) - 1
            while i <= last:
                k = order[i]
                if partner[k] == DELETE and counter[k] >= int(P["minimum_merge_partners"]):   # its whole mass was handed out
                    order[i], order[last] = order[last], order[i]
                    last -= 1
                    continue
                i += 1
            keep = order[:last + 1]
            x, v, m = x[keep], v[keep], m[keep]
    elif P["splitting"]:
        cls = classes()
        nx, nv, nm = [], [], []
        for i in range(len(cls)):
            if cls[i] != 4:
                continue
            nc = min(int(np.floor(m[i] / target(i) + 0.5)), split_patterns.max_children)
            pat = split_patterns.get(nc).astype(np.float64)
            rad = np.sqrt(m[i] / 1.0 / np.pi)
            cm, ov, ox = m[i] / nc, v[i].copy(), x[i].copy()
            for c in range(nc):
                px = ox + pat[c] * rad
                if c == 0:
                    m[i], v[i], x[i] = cm, ov, px
                else:
                    nx.append(px); nv.append(ov); nm.append(cm)
            stats[2] += 1
        if nm:
            x, v, m = np.concatenate([x, np.array(nx)]), np.concatenate([v, np.array(nv)]), np.concatenate([m, np.array(nm)])
    return x, v, m, tuple(stats)


def test_resampling_equals_the_python_restatement(asph, oracle64, default_params, split_patterns):
    """Six full steps of C1 (split on odd steps, merge on even ones, share always): after every physics step the resampling
    phase of the oracle is compared with the restatement above, started from the oracle's own post-physics state."""
    sc = asph.SceneConfig.from_yaml(os.path.join(ROOT, "configs", "default-scene.yaml"))
    params = default_params.replace(max_mass_transfer_sharing=40.0)   # partial hand-overs too, not only whole excesses
    sim = asph.init_fluid_sim(params, sc, split_patterns, lib=oracle64)
    seen = [0, 0, 0]
    for step in range(1, 7):
        dt = sim.single_step_without_adaptivity(params)
        f = {k: sim.get_field(k).astype(np.float64) for k in ("position", "velocity", "mass", "level", "h")}
        off, idx = sim.neighbors_csr()
        rows = [idx[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)]
        x, v, m, stats = _resample(params, split_patterns, f["position"], f["velocity"], f["mass"], f["level"], f["h"], rows, float(dt), step)
        sim.single_step_adaptivity(params, dt)
        info = sim.step_info()
        assert (info["n_shared"], info["n_merged"], info["n_split_parents"]) == stats, (step, info, stats)
        assert sim.num_fluid_particles() == len(m), step
        for name, ref in (("position", x), ("velocity", v), ("mass", m)):
            got = sim.get_field(name).astype(np.float64)
            assert np.abs(got - ref).max() <= 2e-7 * max(np.abs(ref).max(), 1e-12), (step, name)   # float read-back
        seen = [a + b for a, b in zip(seen, stats)]
    assert all(s > 0 for s in seen), seen   # every phase really happened
    sim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["HybridDFSPH", "IISPH"])
def test_cuda_step_equals_the_numpy_restatement(asph, cuda_lib, default_params, solver):
    """The CUDA path against the numpy restatement directly (no oracle in between): one-sweep step in the corner of the
    tank, fp32 tolerances of tests/test_gpu_parity.py."""
    rng = np.random.default_rng(7)
    sp = 0.05
    sc = asph.SceneConfig.dam_break(sp, pos=(-0.999, -0.999), size=(0.6, 0.5))
    pos, vel, mass = asph.scene_particles(sc)
    pos = np.maximum((pos + rng.uniform(-0.15, 0.15, pos.shape) * sp).astype(np.float32), np.float32(-0.999))
    mass = (mass * np.exp(rng.uniform(-0.4, 0.4, mass.shape))).astype(np.float32)
    vel = (rng.standard_normal(vel.shape) * 0.2).astype(np.float32)
    params = default_params.replace(merging=False, sharing=False, splitting=False, level_estimation_method="None", max_iters=0,
                                    pressure_solver_method=solver, hybrid_dfsph_factor=30.0)
    b = asph.scene_boundary(sc, "AnalyticOverestimate")
    probes = [(lambda x, k=k: float(b.planes[k][0]) * x[0] + float(b.planes[k][1]) * x[1] + float(b.planes[k][2])) for k in range(b.n_planes)]
    sim = asph.FluidSimulation(params, pos, vel, mass, b, lib=cuda_lib)
    sim.single_step_without_adaptivity()
    ref = Restatement(cuda_lib, params, pos, vel, mass, probes).step(solver)   # asph_lambda / asph_dlambda: host helpers of the library
    for name in ("density", "aii", "ppe_source_term", "pressure", "pressure_accel"):
        got = sim.get_field(name).astype(np.float64)
        scale = max(np.abs(ref[name]).max(), 1e-12)
        assert np.abs(got - ref[name]).max() <= 2e-4 * scale, (name, np.abs(got - ref[name]).max() / scale)
    assert np.abs(sim.get_field("position") - ref["position"]).max() <= 2e-6
    vs = max(np.abs(ref["velocity"]).max(), 1e-3)
    assert np.abs(sim.get_field("velocity") - ref["velocity"]).max() <= 1e-4 * vs
    sim.close()
