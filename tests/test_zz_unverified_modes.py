"""GPU parity tests of modes whose CUDA kernels were written when the GPU budget of round 1 was (all but) spent.

The library keeps answering ASPH_ERR_UNSUPPORTED for these modes unless ASPH_UNVERIFIED_MODES=1 is set (capi.cu), which
these tests do.  The last GPU seconds of the round went into one run of this file (profiles/r1_w2020_first_hw_run.txt):
the four single-step tests passed on a B200; the multi-step run with resampling failed on a stale error flag of the
list-pool retry path, fixed since but not re-run — that test stays `xfail(strict=False)` (it reports XPASS / xfail without
turning the suite red) and the file is named to run last.  Once it passes the tests move to test_gpu_parity.py and the
switch goes away.

  * operator_discretization: Winchenbach2020 (simulation.rs:1571-1579, boundary_winchenbach2020.rs:207-213, 236-269;
    5 of the reference's media jobs) — k_aii_w2020 (neighbors.cu), k_source / k_sweep<1, ., ., W2020> (solver.cu).
  * support_length_estimation: FromDistribution* and pressure_solver_method: IISPH2 (rest of this file) — never run on hardware.
"""
import numpy as np
import pytest

from test_gpu_parity import _compare_step_fields, _pair, _rel, _scene, _uniform_params

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(90)]
_xf_pending = pytest.mark.xfail(strict=False, reason="failed in its only hardware run (stale flag after a list-pool retry); fixed, re-run pending")
_iso = pytest.mark.isolated(timeout=60)  # body in a child pytest process (tests/conftest.py): a hang or a sticky CUDA error costs this test only


def pending(f):
    return _xf_pending(_iso(f))


@pytest.fixture(autouse=True)
def _enable_unverified(monkeypatch):
    monkeypatch.setenv("ASPH_UNVERIFIED_MODES", "1")


def _one_step(asph, cuda_lib, oracle32, params, pos, vel, mass, boundary):
    g, o = _pair(asph, cuda_lib, oracle32, params, pos, vel, mass, boundary)
    dg = g.single_step_without_adaptivity(); do = o.single_step_without_adaptivity()
    assert dg == do
    gi, oi = g.step_info(), o.step_info()
    assert (gi["div_sweeps"], gi["density_sweeps"]) == (oi["div_sweeps"], oi["density_sweeps"]), (gi, oi)
    pmax = max(float(np.abs(o.get_field("pressure")).max()), 1e-6)
    amax = max(float(np.abs(o.get_field("pressure_accel")).max()), 1e-6)
    w = _compare_step_fields(g, o, 2e-4, [("density", 1.0), ("aii", None), ("ppe_source_term", None), ("pressure", pmax),
                                          ("pressure_accel", amax)])
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= 1e-6, w
    vs = max(float(np.abs(o.get_field("velocity")).max()), 1e-3)
    assert _rel(g.get_field("velocity"), o.get_field("velocity"), vs) <= 1e-4, w
    g.close(); o.close()


@pytest.mark.parametrize("solver", ["HybridDFSPH", "IISPH", "OnlyDivergence"])
def test_winchenbach2020_single_step_uniform(asph, cuda_lib, oracle32, default_params, solver):
    """One physics step, uniform h, block in the corner (boundary terms active), every per-particle field."""
    sc = asph.SceneConfig.dam_break(0.02)
    pos, vel, mass = asph.scene_particles(sc)
    vel = (np.random.default_rng(1).standard_normal(vel.shape) * 0.05).astype(np.float32)
    params = _uniform_params(default_params, pressure_solver_method=solver, operator_discretization="Winchenbach2020")
    _one_step(asph, cuda_lib, oracle32, params, pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))


def test_winchenbach2020_single_step_mixed_sizes(asph, cuda_lib, oracle32, default_params):
    """Jittered cloud with masses spread over 4:1 around the lattice mass (two size levels, far tables), polygon boundary."""
    sc = asph.SceneConfig.dam_break(0.02)
    pos, vel, mass = asph.scene_particles(sc)
    rng = np.random.default_rng(2)
    pos = (pos + rng.uniform(-0.2, 0.2, pos.shape).astype(np.float32) * np.float32(0.02)).astype(np.float32)
    mass = (mass * np.exp(rng.uniform(-np.log(2.0), np.log(2.0), mass.shape))).astype(np.float32)
    vel = (rng.standard_normal(vel.shape) * 0.05).astype(np.float32)
    params = _uniform_params(default_params, operator_discretization="Winchenbach2020", init_boundary_handler="AnalyticUnderestimate")
    _one_step(asph, cuda_lib, oracle32, params, pos, vel, mass, asph.scene_boundary(sc, "AnalyticUnderestimate"))


@pending
def test_winchenbach2020_default_scene_with_resampling(asph, cuda_lib, oracle32, default_params, split_patterns):
    """C1 (default config + scene) under the Winchenbach2020 operator, 15 full steps: identical particle counts and
    resampling statistics every step, positions within 1e-5 of the domain size."""
    sc = _scene(asph, "default-scene.yaml")
    params = default_params.replace(operator_discretization="Winchenbach2020")
    g = asph.init_fluid_sim(params, sc, split_patterns, lib=cuda_lib)
    o = asph.init_fluid_sim(params, sc, split_patterns, lib=oracle32)
    for step in range(15):
        g.single_step(); o.single_step()
        gi, oi = g.step_info(), o.step_info()
        for k in ("n_particles_end", "n_shared", "n_merged", "n_split_parents", "div_sweeps", "density_sweeps"):
            assert gi[k] == oi[k], (step, k, gi, oi)
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= 1e-5
    assert abs(float(g.get_field("mass").sum()) - float(o.get_field("mass").sum())) < 1e-5
    g.close(); o.close()


# ---- support_length_estimation != FromMass (simulation.rs:1873-1971, 1998-2016): k_estimate_h (neighbors.cu), the h2_next /
# lambda state carried through the reorder (grid.cu) and the resampling kernels (adapt.cu).  Never run on hardware.
_xf_never_run = pytest.mark.xfail(strict=False, reason="kernels written after the round's GPU budget was spent: first run on hardware pending")


def never_run(f):
    return _xf_never_run(_iso(f))
H_MODES = ["FromDistribution", "FromDistributionClamped1", "FromDistributionClamped2", "FromDistribution2"]


@never_run
@pytest.mark.parametrize("mode", H_MODES)
def test_support_length_from_distribution_physics(asph, cuda_lib, oracle32, default_params, mode):
    """Four physics steps of C1 with the level set on: h of every step (the previous step's estimate; W summed in list
    order instead of index order, so a few ulp apart), neighbour counts, surface flags, positions."""
    sc = _scene(asph, "default-scene.yaml")
    params = default_params.replace(support_length_estimation=mode, merging=False, sharing=False, splitting=False)
    g = asph.init_fluid_sim(params, sc, None, lib=cuda_lib)
    o = asph.init_fluid_sim(params, sc, None, lib=oracle32)
    for step in range(4):
        dg = g.single_step_without_adaptivity(); do = o.single_step_without_adaptivity()
        assert abs(dg - do) <= 1e-6 * do, step
        hg, ho = g.get_field("h"), o.get_field("h")
        assert np.allclose(hg, ho, rtol=3e-6, atol=0), (step, np.abs(hg / ho - 1).max())
        same = g.get_field("neighbor_count") == o.get_field("neighbor_count")
        assert same.mean() > 0.995, (step, same.mean())  # a pair exactly at the support edge may flip with an ulp of h
        assert (g.get_field("flag_is_fluid_surface") == o.get_field("flag_is_fluid_surface")).mean() > 0.995, step
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= 1e-5
    g.close(); o.close()


@never_run
@pytest.mark.parametrize("mode", ["FromDistributionClamped1", "FromDistribution"])
def test_support_length_from_distribution_with_resampling(asph, cuda_lib, oracle32, default_params, split_patterns, mode):
    """C1 with share / merge / split, 12 full steps: h2_next follows the particles through the resampling kernels —
    same particle counts and resampling statistics every step, masses conserved, positions close."""
    sc = _scene(asph, "default-scene.yaml")
    params = default_params.replace(support_length_estimation=mode)
    g = asph.init_fluid_sim(params, sc, split_patterns, lib=cuda_lib)
    o = asph.init_fluid_sim(params, sc, split_patterns, lib=oracle32)
    for step in range(12):
        g.single_step(); o.single_step()
        gi, oi = g.step_info(), o.step_info()
        for k in ("n_particles_end", "n_shared", "n_merged", "n_split_parents"):
            assert gi[k] == oi[k], (step, k, gi, oi)
    assert abs(float(g.get_field("mass").sum()) - float(o.get_field("mass").sum())) < 1e-5
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= 1e-4
    g.close(); o.close()


# ---- pressure_solver_method: IISPH2 (simulation.rs:2262-2387): k_omega, the omega-scaled source, k_scale_pressure (solver.cu),
# the size classes of the last resampling phase carried to the next step (grid.cu, adapt.cu).  Never run on hardware.
@never_run
def test_iisph2_single_step_uniform(asph, cuda_lib, oracle32, default_params):
    sc = asph.SceneConfig.dam_break(0.02)
    pos, vel, mass = asph.scene_particles(sc)
    vel = (np.random.default_rng(1).standard_normal(vel.shape) * 0.05).astype(np.float32)
    params = _uniform_params(default_params, pressure_solver_method="IISPH2")
    _one_step(asph, cuda_lib, oracle32, params, pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))


@never_run
def test_iisph2_default_scene_with_resampling(asph, cuda_lib, oracle32, default_params, split_patterns):
    """C1, 12 full steps: the Large particles of each resampling phase take the single-term omega in the next step."""
    sc = _scene(asph, "default-scene.yaml")
    params = default_params.replace(pressure_solver_method="IISPH2")
    g = asph.init_fluid_sim(params, sc, split_patterns, lib=cuda_lib)
    o = asph.init_fluid_sim(params, sc, split_patterns, lib=oracle32)
    for step in range(12):
        g.single_step(); o.single_step()
        gi, oi = g.step_info(), o.step_info()
        for k in ("n_particles_end", "n_shared", "n_merged", "n_split_parents", "density_sweeps"):
            assert gi[k] == oi[k], (step, k, gi, oi)
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= 1e-5
    g.close(); o.close()


# ---- experiment ASPH_ROWS4=1 (not a mode of the reference: a faster row schedule for the sweep kernels):
# the neighbour pass writes the particle's own row last, k_sweep<., ., ., ., R4> stops one row early and in steps of 4 rows.
# Never run on hardware; results must stay within the tolerances of the default schedule.
@pytest.fixture
def rows4(monkeypatch):
    monkeypatch.setenv("ASPH_ROWS4", "1")
    monkeypatch.delenv("ASPH_UNVERIFIED_MODES", raising=False)


@never_run
def test_rows4_neighbor_sets_bit_exact(asph, cuda_lib, oracle32, default_params, rows4):
    from test_gpu_parity import _point_clouds
    b = asph.scene_boundary(_scene(asph, "default-scene.yaml"), "AnalyticOverestimate")
    for name, (pos, mass) in _point_clouds(asph).items():
        g, o = _pair(asph, cuda_lib, oracle32, default_params, pos, np.zeros_like(pos), mass, b)
        g.build_neighbors(np.float32(2.0)); o.build_neighbors(np.float32(2.0))
        go, gi = g.neighbors_csr(); oo, oi = o.neighbors_csr()
        assert np.array_equal(go, oo) and np.array_equal(gi, oi), name
        g.close(); o.close()


@never_run
@pytest.mark.parametrize("solver", ["HybridDFSPH", "IISPH"])
def test_rows4_single_step_uniform(asph, cuda_lib, oracle32, default_params, solver, rows4):
    from test_gpu_parity import _single_step_uniform
    _single_step_uniform(asph, cuda_lib, oracle32, default_params, solver)


@never_run
def test_rows4_single_step_mixed_sizes(asph, cuda_lib, oracle32, default_params, rows4):
    sc = asph.SceneConfig.dam_break(0.02)
    pos, vel, mass = asph.scene_particles(sc)
    rng = np.random.default_rng(2)
    pos = (pos + rng.uniform(-0.2, 0.2, pos.shape).astype(np.float32) * np.float32(0.02)).astype(np.float32)
    mass = (mass * np.exp(rng.uniform(-np.log(2.0), np.log(2.0), mass.shape))).astype(np.float32)
    vel = (rng.standard_normal(vel.shape) * 0.05).astype(np.float32)
    _one_step(asph, cuda_lib, oracle32, _uniform_params(default_params), pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))


@never_run
def test_rows4_default_scene_with_resampling(asph, cuda_lib, oracle32, default_params, split_patterns, rows4):
    sc = _scene(asph, "default-scene.yaml")
    g = asph.init_fluid_sim(default_params, sc, split_patterns, lib=cuda_lib)
    o = asph.init_fluid_sim(default_params, sc, split_patterns, lib=oracle32)
    for step in range(20):
        g.single_step(); o.single_step()
        gi, oi = g.step_info(), o.step_info()
        for k in ("n_particles_end", "n_shared", "n_merged", "n_split_parents", "div_sweeps", "density_sweeps"):
            assert gi[k] == oi[k], (step, k, gi, oi)
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= 1e-5
    g.close(); o.close()


@never_run
def test_rows4_two_gpu_pressure_solve_matches_single_gpu(monkeypatch):
    """The PEER instantiations of the 4-row sweep kernels (needs 2 GPUs; the single-GPU reference run inside the worker
    uses the 4-row kernels as well)."""
    import test_dist_gpu as D
    if D._gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("ASPH_ROWS4", "1")
    rep = D._run(2, 4, "HybridDFSPH", mode="random")
    D._check(rep, 1e-6)
    assert rep["sweeps_equal"], rep


# ---- experiment ASPH_BULK=1: interior tiles of the sweep kernels staged by cp.async.bulk + mbarrier
# (k_sweep_bulk, sweep_kernel.inc).  Never run on hardware; results must equal the default kernels' (same arithmetic, same order).
@pytest.fixture
def bulk(monkeypatch):
    monkeypatch.setenv("ASPH_BULK", "1")
    monkeypatch.delenv("ASPH_UNVERIFIED_MODES", raising=False)
    monkeypatch.delenv("ASPH_ROWS4", raising=False)


@never_run
@pytest.mark.parametrize("solver", ["HybridDFSPH", "IISPH"])
def test_bulk_single_step_uniform(asph, cuda_lib, oracle32, default_params, solver, bulk):
    from test_gpu_parity import _single_step_uniform
    _single_step_uniform(asph, cuda_lib, oracle32, default_params, solver)


@never_run
def test_bulk_single_step_mixed_sizes(asph, cuda_lib, oracle32, default_params, bulk):
    """{h, m} window variant: masses spread over 4:1."""
    sc = asph.SceneConfig.dam_break(0.02)
    pos, vel, mass = asph.scene_particles(sc)
    rng = np.random.default_rng(2)
    pos = (pos + rng.uniform(-0.2, 0.2, pos.shape).astype(np.float32) * np.float32(0.02)).astype(np.float32)
    mass = (mass * np.exp(rng.uniform(-np.log(2.0), np.log(2.0), mass.shape))).astype(np.float32)
    vel = (rng.standard_normal(vel.shape) * 0.05).astype(np.float32)
    _one_step(asph, cuda_lib, oracle32, _uniform_params(default_params), pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))


@never_run
def test_bulk_equals_default_kernels_bit_for_bit(asph, cuda_lib, default_params, monkeypatch):
    """Same arithmetic in the same order: 10 steps of a 78 k-particle uniform dam break (about 300 interior tiles) end in
    the bit-identical state with and without the switch."""
    sc = asph.SceneConfig.dam_break(0.004)
    pos, vel, mass = asph.scene_particles(sc)
    vel = (np.random.default_rng(3).standard_normal(vel.shape) * 0.05).astype(np.float32)
    params = _uniform_params(default_params)
    b = asph.scene_boundary(sc, "AnalyticOverestimate")
    out = []
    for flag in ("0", "1"):
        monkeypatch.setenv("ASPH_BULK", flag)
        g = asph.FluidSimulation(params, pos, vel, mass, b, lib=cuda_lib)
        sweeps = []
        for _ in range(10):
            g.single_step(); i = g.step_info(); sweeps.append((i["div_sweeps"], i["density_sweeps"]))
        out.append((sweeps, g.get_field("position"), g.get_field("velocity"), g.get_field("pressure")))
        g.close()
    assert out[0][0] == out[1][0]
    for k in (1, 2, 3):
        assert np.array_equal(out[0][k], out[1][k]), k


@never_run
def test_rows4_with_bulk_default_scene_with_resampling(asph, cuda_lib, oracle32, default_params, split_patterns, monkeypatch):
    """Both experiments together (k_sweep_bulk<., ., ., ., R4>): C1 with level set + resampling, 12 steps."""
    monkeypatch.setenv("ASPH_ROWS4", "1"); monkeypatch.setenv("ASPH_BULK", "1")
    monkeypatch.delenv("ASPH_UNVERIFIED_MODES", raising=False)
    sc = _scene(asph, "default-scene.yaml")
    g = asph.init_fluid_sim(default_params, sc, split_patterns, lib=cuda_lib)
    o = asph.init_fluid_sim(default_params, sc, split_patterns, lib=oracle32)
    for step in range(12):
        g.single_step(); o.single_step()
        gi, oi = g.step_info(), o.step_info()
        for k in ("n_particles_end", "n_shared", "n_merged", "n_split_parents", "div_sweeps", "density_sweeps"):
            assert gi[k] == oi[k], (step, k, gi, oi)
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= 1e-5
    g.close(); o.close()

