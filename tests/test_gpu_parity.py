"""GPU parity: the CUDA step (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): neighbour indices / counts bit-exact; floating-point fields within a stated fp32
tolerance.  The oracle sums neighbours in ascending reference index, the GPU in grid-cell order, so pair sums differ
by fp32 rounding (a few ulp of the largest term); tolerances below are relative to the field's scale.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene(asph, name):
    return asph.SceneConfig.from_yaml(os.path.join(ROOT, "configs", name))


def _pair(asph, cuda_lib, oracle, params, pos, vel, mass, boundary, split=None):
    g = asph.FluidSimulation(params, pos, vel, mass, boundary, split, lib=cuda_lib)
    o = asph.FluidSimulation(params, pos, vel, mass, boundary, split, lib=oracle)
    return g, o


def _uniform_params(default_params, **kw):
    base = dict(merging=False, sharing=False, splitting=False, level_estimation_method="None")
    base.update(kw)
    return default_params.replace(**base)


def _rel(a, b, scale=None):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    s = scale if scale is not None else max(np.abs(b).max(), 1e-30)
    return np.abs(a - b).max() / s


def _point_clouds(asph):
    rng = np.random.default_rng(0)
    out = {}
    sc = _scene(asph, "default-scene.yaml")
    pos, vel, mass = asph.scene_particles(sc)
    out["default-scene (two spacings)"] = (pos, mass)
    sp = np.float32(0.01)
    blk = dict(pos=(np.float32(-0.5), np.float32(-0.5)), size=(np.float32(0.6), np.float32(0.4)), spacing=sp,
               volume_fill_ratio=np.float32(0.93), velocity=(np.float32(0), np.float32(0)))
    p, _, m = asph.add_fluid_block(blk)
    out["lattice 60x40 (exact ties)"] = (p, m)
    pj = (p + rng.uniform(-0.3, 0.3, p.shape).astype(np.float32) * sp).astype(np.float32)
    out["jittered lattice"] = (pj, m)
    # strongly mixed sizes: masses spread over 16^2 : 1
    mm = (m * np.exp(rng.uniform(0, np.log(256.0), m.shape))).astype(np.float32)
    out["jittered, mass ratio 256"] = (pj, mm)
    out["single particle"] = (p[:1].copy(), m[:1].copy())
    out["two coincident particles"] = (np.repeat(p[:1], 2, axis=0), m[:2].copy())
    return out


@pytest.mark.parametrize("f", [2.0, 5.5 / 1.9])
def test_neighbor_sets_bit_exact(asph, cuda_lib, oracle32, default_params, f):
    """N_f(i) identical (indices and counts) to the oracle's grid search, which test_oracle_neighbors pins to the
    reference's brute-force predicate (simulation.rs:1810-1863)."""
    b = asph.scene_boundary(_scene(asph, "default-scene.yaml"), "AnalyticOverestimate")
    for name, (pos, mass) in _point_clouds(asph).items():
        vel = np.zeros_like(pos)
        g, o = _pair(asph, cuda_lib, oracle32, default_params, pos, vel, mass, b)
        g.build_neighbors(np.float32(f)); o.build_neighbors(np.float32(f))
        go, gi = g.neighbors_csr(); oo, oi = o.neighbors_csr()
        assert np.array_equal(go, oo), name
        assert np.array_equal(gi, oi), name
        assert np.array_equal(g.get_field("neighbor_count"), o.get_field("neighbor_count")), name
        assert np.array_equal(g.get_field("h"), o.get_field("h")), name  # h bit-exact
        g.close(); o.close()


def _compare_step_fields(g, o, tol, fields):
    worst = {}
    for name, scale in fields:
        a, b = g.get_field(name), o.get_field(name)
        worst[name] = _rel(a, b, scale)
    bad = {k: v for k, v in worst.items() if not (v <= tol)}
    assert not bad, f"fields beyond tolerance {tol}: {bad}; all: {worst}"
    return worst


def _single_step_uniform(asph, cuda_lib, oracle32, default_params, solver, spacing=0.02, vel_scale=0.05,
                         boundary_kind="AnalyticOverestimate", block_pos=(-0.95, -0.9)):
    """One physics step, uniform h, level estimation off (the reference's 'Uniform SPH' recipe,
    media/motivation-video.yaml:42-57): every per-particle quantity against the oracle."""
    sc = asph.SceneConfig.dam_break(spacing, pos=block_pos)
    pos, vel, mass = asph.scene_particles(sc)
    rng = np.random.default_rng(1)
    vel = (rng.standard_normal(vel.shape) * vel_scale).astype(np.float32)
    params = _uniform_params(default_params, pressure_solver_method=solver, init_boundary_handler=boundary_kind)
    b = asph.scene_boundary(sc, boundary_kind)
    g, o = _pair(asph, cuda_lib, oracle32, params, pos, vel, mass, b)
    dg = g.single_step_without_adaptivity(); do = o.single_step_without_adaptivity()
    assert dg == do  # dt is an exact min-reduction
    gi, oi = g.step_info(), o.step_info()
    assert (gi["div_sweeps"], gi["density_sweeps"]) == (oi["div_sweeps"], oi["density_sweeps"]), (gi, oi)
    assert np.array_equal(g.get_field("neighbor_count"), o.get_field("neighbor_count"))
    # boundary terms: same operation order; the two λ tables agree to 1e-8 in double, i.e. to an fp32 ulp
    assert np.allclose(g.get_field("lambda_sum"), o.get_field("lambda_sum"), rtol=1e-6, atol=1e-7)
    assert np.allclose(g.get_field("lambda_grad"), o.get_field("lambda_grad"), rtol=1e-6, atol=1e-5)
    pmax = max(float(np.abs(o.get_field("pressure")).max()), 1e-6)
    amax = max(float(np.abs(o.get_field("pressure_accel")).max()), 1e-6)
    w = _compare_step_fields(g, o, 2e-4, [("density", 1.0), ("aii", None), ("ppe_source_term", None), ("pressure", pmax),
                                          ("pressure_accel", amax)])
    # positions / velocities: 1e-5 relative to the domain size (2 m) resp. the velocity scale
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= 1e-6, w
    vs = max(float(np.abs(o.get_field("velocity")).max()), 1e-3)
    assert _rel(g.get_field("velocity"), o.get_field("velocity"), vs) <= 1e-4, w
    g.close(); o.close()



@pytest.mark.parametrize("solver", ["HybridDFSPH", "IISPH", "OnlyDivergence"])
def test_single_step_uniform(asph, cuda_lib, oracle32, default_params, solver):
    _single_step_uniform(asph, cuda_lib, oracle32, default_params, solver)


@pytest.mark.parametrize("solver", ["HybridDFSPH", "IISPH"])
def test_single_step_uniform_many_tiles_per_block(asph, cuda_lib, oracle32, default_params, solver, monkeypatch):
    """The same step with the persistent sweep kernels limited to two blocks (ASPH_SWEEP_GRID, a test hook of
    solver.cu), so that each block walks ~50 tiles through its two-stage copy pipeline."""
    monkeypatch.setenv("ASPH_SWEEP_GRID", "2")
    _single_step_uniform(asph, cuda_lib, oracle32, default_params, solver, spacing=0.01)


@pytest.mark.parametrize("solver", ["HybridDFSPH", "IISPH"])
def test_single_step_polygon_boundary(asph, cuda_lib, oracle32, default_params, solver):
    """`init_boundary_handler: AnalyticUnderestimate` (sdf/sdf2d.rs: one polygon SDF instead of four planes; used by
    media/video-viscosity.yaml): the block sits in the corner, within the support radius of the floor and the wall, so
    edges and the corner vertex both decide λ, ∇λ and through them every other field."""
    _single_step_uniform(asph, cuda_lib, oracle32, default_params, solver, spacing=0.02, boundary_kind="AnalyticUnderestimate",
                         block_pos=(-0.985, -0.985))


def test_trajectory_polygon_boundary_with_resampling(asph, cuda_lib, oracle32, default_params, split_patterns):
    """C1 (default-config + default-scene, level set and resampling on) with the polygon boundary, 20 steps."""
    sc = _scene(asph, "default-scene.yaml")
    params = default_params.replace(init_boundary_handler="AnalyticUnderestimate")
    g = asph.init_fluid_sim(params, sc, split_patterns, lib=cuda_lib)
    o = asph.init_fluid_sim(params, sc, split_patterns, lib=oracle32)
    for k in range(20):
        g.single_step(); o.single_step()
        assert g.num_fluid_particles() == o.num_fluid_particles(), k
    err = np.abs(g.get_field("position") - o.get_field("position")).max() / 2.0
    print(f"C1 polygon boundary, 20 steps: N = {g.num_fluid_particles()}, |gpu - oracle32| / L = {err:.3e}")
    assert err <= 1e-5
    g.close(); o.close()


def test_full_size_step_against_oracle(asph, cuda_lib, oracle32, default_params):
    """BASELINE config[1] at full size (999 292 particles): one HybridDFSPH step with seeded random velocities (both
    solves iterate) against the CPU oracle — the far tables, 32-column strips and multi-tile pipelines at scale."""
    _single_step_uniform(asph, cuda_lib, oracle32, default_params, "HybridDFSPH", spacing=1.122e-3, vel_scale=0.02)


def test_level_estimation_default_config(asph, cuda_lib, oracle32, default_params):
    """C1 (default-config + default-scene): surface detection, level-set propagation and smoothing."""
    sc = _scene(asph, "default-scene.yaml")
    pos, vel, mass = asph.scene_particles(sc)
    params = default_params
    b = asph.scene_boundary(sc, "AnalyticOverestimate")
    g, o = _pair(asph, cuda_lib, oracle32, params, pos, vel, mass, b)
    g.single_step_without_adaptivity(); o.single_step_without_adaptivity()
    assert np.array_equal(g.get_field("flag_is_fluid_surface"), o.get_field("flag_is_fluid_surface"))
    assert np.array_equal(g.get_field("flag_insufficient_neighs"), o.get_field("flag_insufficient_neighs"))
    assert g.step_info()["level_sweeps"] == o.step_info()["level_sweeps"]
    lg, lo = g.get_field("level"), o.get_field("level")
    assert np.abs(lg - lo).max() <= 1e-5 * max(1.0, np.abs(lo).max()), np.abs(lg - lo).max()
    g.close(); o.close()


def test_trajectory_c1_physics(asph, cuda_lib, oracle32, oracle64, default_params):
    """20 physics steps of C1 without resampling: GPU vs fp32 oracle, judged against the fp32-vs-fp64 oracle
    distance (the noise floor of the reference's own arithmetic)."""
    sc = _scene(asph, "default-scene.yaml")
    pos, vel, mass = asph.scene_particles(sc)
    params = default_params.replace(merging=False, sharing=False, splitting=False)
    b = asph.scene_boundary(sc, "AnalyticOverestimate")
    g = asph.FluidSimulation(params, pos, vel, mass, b, lib=cuda_lib)
    o = asph.FluidSimulation(params, pos, vel, mass, b, lib=oracle32)
    d = asph.FluidSimulation(params, pos, vel, mass, b, lib=oracle64)
    for _ in range(20):
        g.single_step(); o.single_step(); d.single_step()
    pg, po, pd = g.get_field("position"), o.get_field("position"), d.get_field("position")
    err_gpu = np.abs(pg - po).max() / 2.0
    floor = np.abs(po - pd).max() / 2.0
    print(f"C1 20 steps: |gpu - oracle32| / L = {err_gpu:.3e}; |oracle32 - oracle64| / L = {floor:.3e}")
    assert err_gpu <= max(1e-5, 3 * floor)
    g.close(); o.close(); d.close()


def test_full_size_properties(asph, cuda_lib, default_params):
    """BASELINE config[1] (999 292 particles, uniform h, HybridDFSPH) through size-independent properties:
    neighbour lists symmetric with self included, lattice-interior density ~ rho0 neighbours, finite state,
    mass untouched, a_ii equals the operator diagonal on sampled particles (check_aii, simulation.rs:1324-1375)."""
    sc = asph.SceneConfig.dam_break(1.122e-3)
    pos, vel, mass = asph.scene_particles(sc)
    assert len(mass) == 999292
    params = _uniform_params(default_params)
    b = asph.scene_boundary(sc, "AnalyticOverestimate")
    g = asph.FluidSimulation(params, pos, vel, mass, b, lib=cuda_lib)
    g.single_step_without_adaptivity()
    off, idx = g.neighbors_csr()
    n = len(mass)
    cnt = np.diff(off).astype(np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), cnt)
    # self included
    assert np.all(np.bincount(rows[idx == rows], minlength=n) == 1)
    # symmetric: the multiset of (i, j) equals the multiset of (j, i)
    a = rows * n + idx
    bb = idx.astype(np.int64) * n + rows
    assert np.array_equal(np.sort(a), np.sort(bb))
    assert np.all(np.isfinite(g.get_field("position"))) and np.all(np.isfinite(g.get_field("velocity")))
    assert np.array_equal(g.get_field("mass"), mass)
    rho = g.get_field("density")
    interior = cnt == cnt.max()
    # lattice at rest: density = volume_fill_ratio * rho0 up to the kernel's lattice quadrature error
    assert abs(float(rho[interior].mean()) - 0.93) < 0.02
    g.close()


# ------------------------------------------------------------------------------------------------ resampling
def _adaptive_case(asph, default_params, seed, n_side=48):
    """A jittered lattice with masses spread over all five size classes of a prescribed level field."""
    rng = np.random.default_rng(seed)
    sp = np.float32(0.01)
    blk = dict(pos=(np.float32(-0.4), np.float32(-0.4)), size=(np.float32(n_side * 0.01 + 0.001), np.float32(n_side * 0.01 + 0.001)),
               spacing=sp, volume_fill_ratio=np.float32(0.93), velocity=(np.float32(0), np.float32(0)))
    pos, vel, mass = asph.add_fluid_block(blk)
    pos = (pos + rng.uniform(-0.25, 0.25, pos.shape).astype(np.float32) * sp).astype(np.float32)
    vel = (rng.standard_normal(pos.shape) * 0.1).astype(np.float32)
    mass = (mass * rng.uniform(0.3, 2.6, mass.shape)).astype(np.float32)
    level = (-(0.08 - pos[:, 1]) * 0.5).astype(np.float32)          # "depth" below y = 0.08, always <= 0 here
    level = np.minimum(level, np.float32(0.0))
    r0 = float(np.sqrt(0.93e-4 / np.pi))
    params = default_params.replace(particle_radius_fine=r0, particle_radius_base=2.0 * r0, maximum_surface_distance=0.2)
    return pos, vel, mass, level, params


def _hooks(g, o):
    import ctypes as C
    fp = C.POINTER(C.c_float)
    g.lib.asph_set_level.argtypes = [C.c_void_p, fp, C.c_uint64]; g.lib.asph_set_level.restype = C.c_int
    g.lib.asph_set_step_number.argtypes = [C.c_void_p, C.c_uint64]; g.lib.asph_set_step_number.restype = None
    o.lib.oracle_set_level.argtypes = [C.c_void_p, fp, C.c_uint64]; o.lib.oracle_set_level.restype = C.c_int
    o.lib.oracle_set_step_number.argtypes = [C.c_void_p, C.c_uint64]; o.lib.oracle_set_step_number.restype = None


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("phase", ["share", "merge", "split", "share+merge", "share+split"])
def test_resampling_phase_parity(asph, cuda_lib, oracle32, default_params, split_patterns, phase, seed):
    """single_step_adaptivity on identical inputs (state, 2h neighbour lists, level field): the same donors pick the
    same receivers, the same particles are deleted / split, and the resulting particle set is BIT-identical, in
    the reference's particle order (swap-with-last deletion, append-at-end children)."""
    import ctypes as C
    pos, vel, mass, level, params = _adaptive_case(asph, default_params, seed)
    params = params.replace(sharing="share" in phase, merging="merge" in phase, splitting="split" in phase)
    step_number = 2 if "merge" in phase else 3
    b = asph.scene_boundary(_scene(asph, "default-scene.yaml"), "AnalyticOverestimate")
    g, o = _pair(asph, cuda_lib, oracle32, params, pos, vel, mass, b, split_patterns)
    _hooks(g, o)
    g.build_neighbors(np.float32(2.0)); o.build_neighbors(np.float32(2.0))
    lp = level.ctypes.data_as(C.POINTER(C.c_float))
    assert g.lib.asph_set_level(g._h, lp, len(level)) == 0
    assert o.lib.oracle_set_level(o._h, lp, len(level)) == 0
    g.lib.asph_set_step_number(g._h, step_number); o.lib.oracle_set_step_number(o._h, step_number)
    dt = 0.002
    g.single_step_adaptivity(dt=dt); o.single_step_adaptivity(dt=dt)
    gi, oi = g.step_info(), o.step_info()
    assert (gi["n_shared"], gi["n_merged"], gi["n_split_parents"]) == (oi["n_shared"], oi["n_merged"], oi["n_split_parents"]), (gi, oi)
    if "share" in phase:
        assert oi["n_shared"] > 0
    if "merge" in phase:
        assert oi["n_merged"] > 0
    if "split" in phase:
        assert oi["n_split_parents"] > 0
    assert g.num_fluid_particles() == o.num_fluid_particles()
    for f in ("mass", "position", "velocity"):
        assert np.array_equal(g.get_field(f), o.get_field(f)), f
    if phase == "share":
        assert np.array_equal(g.get_field("merge_partner"), o.get_field("merge_partner"))
        assert np.array_equal(g.get_field("merge_counter"), o.get_field("merge_counter"))
        assert np.array_equal(g.get_field("particle_size_class"), o.get_field("particle_size_class"))
    g.close(); o.close()


def test_trajectory_c1_resampling(asph, cuda_lib, oracle32, default_params, split_patterns):
    """C1 with level set + share / merge / split on: particle counts per step and final positions vs the oracle."""
    sc = _scene(asph, "default-scene.yaml")
    g = asph.init_fluid_sim(default_params, sc, split_patterns, lib=cuda_lib)
    o = asph.init_fluid_sim(default_params, sc, split_patterns, lib=oracle32)
    counts = []
    for k in range(40):
        g.single_step(); o.single_step()
        counts.append((g.num_fluid_particles(), o.num_fluid_particles()))
    assert all(a == b for a, b in counts), counts
    err = np.abs(g.get_field("position") - o.get_field("position")).max() / 2.0
    print(f"C1 40 steps with resampling: N = {counts[-1][0]}, |gpu - oracle32| / L = {err:.3e}")
    assert err <= 1e-5
    assert abs(float(g.get_field("mass").sum()) - float(o.get_field("mass").sum())) < 1e-5
    g.close(); o.close()


def test_adaptive_dam_break_mid_size(asph, cuda_lib, oracle32, default_params, split_patterns):
    """BASELINE config[2] recipe at a size the oracle steps in seconds: dam-break block at spacing 0.004 (77 k particles),
    particle_radius_fine = the lattice particle's radius, base radius 4x, maximum_surface_distance 0.2, level set and
    share / merge / split on.  Several size levels, wide slices, far tables and the {h, m} window at scale: particle
    counts identical every step, positions within the north-star tolerance, mass conserved."""
    spacing = 0.004
    sc = asph.SceneConfig.dam_break(spacing)
    r_f = float(np.sqrt(0.93 / np.pi) * spacing)
    params = default_params.replace(particle_radius_fine=r_f, particle_radius_base=4.0 * r_f, maximum_surface_distance=0.2)
    g = asph.init_fluid_sim(params, sc, split_patterns, lib=cuda_lib)
    o = asph.init_fluid_sim(params, sc, split_patterns, lib=oracle32)
    m0 = float(o.get_field("mass").astype(np.float64).sum())
    counts, merged = [], 0
    for k in range(12):
        g.single_step(); o.single_step()
        counts.append((g.num_fluid_particles(), o.num_fluid_particles()))
        gi, oi = g.step_info(), o.step_info()
        assert (gi["n_shared"], gi["n_merged"], gi["n_split_parents"]) == (oi["n_shared"], oi["n_merged"], oi["n_split_parents"]), (k, gi, oi)
        merged += gi["n_merged"]
    assert all(a == b for a, b in counts), counts
    assert merged > 0 and counts[-1][0] < counts[0][0]   # the interior really coarsened
    err = np.abs(g.get_field("position") - o.get_field("position")).max() / 2.0
    print(f"adaptive dam break: N {counts[0][0]} -> {counts[-1][0]}, |gpu - oracle32| / L = {err:.3e}")
    assert err <= 1e-5
    assert abs(float(g.get_field("mass").astype(np.float64).sum()) - m0) < 0.005   # simulation.rs:2791
    ms = g.get_field("mass")
    assert ms.max() / ms.min() > 2.25                     # h ~ sqrt(m): more than one size level in play
    g.close(); o.close()
