"""GPU parity of the modes beside the default configuration (SURVEY.md §8f rank 3 and the parameter variants of §8a):
the CUDA step through the C ABI against the CPU oracle on the same inputs, same tolerances as tests/test_gpu_parity.py.

  * operator_discretization: Winchenbach2020 (simulation.rs:1571-1579, boundary_winchenbach2020.rs:207-213, 236-269),
    ConsistentSymmetricGradient (boundary_winchenbach2020.rs:177-186)
  * support_length_estimation: FromDistribution* (simulation.rs:1873-1971, 1998-2016)
  * pressure_solver_method: IISPH2 (simulation.rs:2262-2387)
  * viscosity_type: WCSPH (simulation.rs:946-966), pull_fluid_to (:991-1003)
  * hybrid_dfsph_density_source_term: OnlyDensity, hybrid_dfsph_non_pressure_accel_before_divergence_free: false
  * boundary_penalty_term: None / Linear / Quadratic2 (boundary_winchenbach2020.rs:87-118)
  * sizing_function: Mass / Radius2 (simulation.rs:213-237)
  * BASELINE configs[3] geometry (ratio-stress-test scene, radius ratio 16:1, IISPH) at a size the oracle steps in seconds
  * constrain_neighborhood_count (simulation.rs:2145-2177), level_estimation_after_advection (:2678-2707), CenterDiff (:631-695)
  * the bulk-copy sweep kernels against the per-thread-copy ones, bit for bit
  * north star: positions within 1e-5 relative after 100 steps (a trajectory that is not chaotic, noise floor printed)
"""
import ctypes as C
import os

import numpy as np
import pytest

from test_gpu_parity import _adaptive_case, _compare_step_fields, _hooks, _pair, _rel, _scene, _uniform_params

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]


def _one_step(asph, cuda_lib, oracle32, params, pos, vel, mass, boundary, sweeps_exact=True):
    g, o = _pair(asph, cuda_lib, oracle32, params, pos, vel, mass, boundary)
    dg = g.single_step_without_adaptivity(); do = o.single_step_without_adaptivity()
    assert dg == do
    gi, oi = g.step_info(), o.step_info()
    if sweeps_exact:
        assert (gi["div_sweeps"], gi["density_sweeps"]) == (oi["div_sweeps"], oi["density_sweeps"]), (gi, oi)
    pmax = max(float(np.abs(o.get_field("pressure")).max()), 1e-6)
    amax = max(float(np.abs(o.get_field("pressure_accel")).max()), 1e-6)
    w = _compare_step_fields(g, o, 2e-4, [("density", 1.0), ("aii", None), ("ppe_source_term", None), ("pressure", pmax),
                                          ("pressure_accel", amax)])
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= 1e-6, w
    vs = max(float(np.abs(o.get_field("velocity")).max()), 1e-3)
    assert _rel(g.get_field("velocity"), o.get_field("velocity"), vs) <= 1e-4, w
    g.close(); o.close()


def _corner_block(asph, spacing=0.02, seed=1, vel_scale=0.05, pos=(-0.95, -0.9)):
    sc = asph.SceneConfig.dam_break(spacing, pos=pos)
    x, v, m = asph.scene_particles(sc)
    v = (np.random.default_rng(seed).standard_normal(v.shape) * vel_scale).astype(np.float32)
    return sc, x, v, m


def _mixed_cloud(asph, seed=2):
    """Jittered cloud with masses spread over 4:1 around the lattice mass (two size levels, far tables)."""
    sc = asph.SceneConfig.dam_break(0.02)
    pos, vel, mass = asph.scene_particles(sc)
    rng = np.random.default_rng(seed)
    pos = (pos + rng.uniform(-0.2, 0.2, pos.shape).astype(np.float32) * np.float32(0.02)).astype(np.float32)
    mass = (mass * np.exp(rng.uniform(-np.log(2.0), np.log(2.0), mass.shape))).astype(np.float32)
    vel = (rng.standard_normal(vel.shape) * 0.05).astype(np.float32)
    return sc, pos, vel, mass


def _steps_with_resampling(asph, cuda_lib, oracle32, params, split_patterns, steps, keys, tol):
    sc = _scene(asph, "default-scene.yaml")
    g = asph.init_fluid_sim(params, sc, split_patterns, lib=cuda_lib)
    o = asph.init_fluid_sim(params, sc, split_patterns, lib=oracle32)
    for step in range(steps):
        g.single_step(); o.single_step()
        gi, oi = g.step_info(), o.step_info()
        for k in keys:
            assert gi[k] == oi[k], (step, k, gi, oi)
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= tol
    assert abs(float(g.get_field("mass").sum()) - float(o.get_field("mass").sum())) < 1e-5
    g.close(); o.close()


RESAMPLING_KEYS = ("n_particles_end", "n_shared", "n_merged", "n_split_parents")


# ---- operator_discretization -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver", ["HybridDFSPH", "IISPH", "OnlyDivergence"])
def test_winchenbach2020_single_step_uniform(asph, cuda_lib, oracle32, default_params, solver):
    """One physics step, uniform h, block in the corner (boundary terms active), every per-particle field."""
    sc, pos, vel, mass = _corner_block(asph)
    params = _uniform_params(default_params, pressure_solver_method=solver, operator_discretization="Winchenbach2020")
    _one_step(asph, cuda_lib, oracle32, params, pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))


def test_winchenbach2020_single_step_mixed_sizes(asph, cuda_lib, oracle32, default_params):
    sc, pos, vel, mass = _mixed_cloud(asph)
    params = _uniform_params(default_params, operator_discretization="Winchenbach2020", init_boundary_handler="AnalyticUnderestimate")
    _one_step(asph, cuda_lib, oracle32, params, pos, vel, mass, asph.scene_boundary(sc, "AnalyticUnderestimate"))


def test_winchenbach2020_default_scene_with_resampling(asph, cuda_lib, oracle32, default_params, split_patterns):
    """C1 under the Winchenbach2020 operator, 15 full steps: identical particle counts, resampling statistics and sweep
    counts every step, positions within 1e-5 of the domain size."""
    _steps_with_resampling(asph, cuda_lib, oracle32, default_params.replace(operator_discretization="Winchenbach2020"), split_patterns, 15,
                           RESAMPLING_KEYS + ("div_sweeps", "density_sweeps"), 1e-5)


@pytest.mark.parametrize("solver", ["HybridDFSPH", "IISPH"])
def test_consistent_symmetric_gradient_single_step(asph, cuda_lib, oracle32, default_params, solver):
    """p_ib = p_i mirrored into the boundary (boundary_winchenbach2020.rs:177-186, :270-304): a_ii and a^p near the walls."""
    sc, pos, vel, mass = _corner_block(asph, pos=(-0.985, -0.985))
    params = _uniform_params(default_params, pressure_solver_method=solver, operator_discretization="ConsistentSymmetricGradient")
    _one_step(asph, cuda_lib, oracle32, params, pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))


# ---- non-pressure forces and sources ---------------------------------------------------------------------------------
@pytest.mark.parametrize("cloud", ["uniform", "mixed"])
def test_wcsph_viscosity_single_step(asph, cuda_lib, oracle32, default_params, cloud):
    """viscosity_type: WCSPH (simulation.rs:946-966; speed of sound 88, 0.001 h^2 regularisation), strong enough to matter."""
    sc, pos, vel, mass = _corner_block(asph, vel_scale=0.3) if cloud == "uniform" else _mixed_cloud(asph)
    params = _uniform_params(default_params, viscosity_type="WCSPH", viscosity=0.05)
    g, o = _pair(asph, cuda_lib, oracle32, params.replace(viscosity=0.0), pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))
    o.single_step_without_adaptivity()
    v_inviscid = o.get_field("velocity").copy()
    g.close(); o.close()
    _one_step(asph, cuda_lib, oracle32, params, pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))
    g, o = _pair(asph, cuda_lib, oracle32, params, pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))
    o.single_step_without_adaptivity()
    assert np.abs(o.get_field("velocity") - v_inviscid).max() > 1e-3  # the term is really exercised
    g.close(); o.close()


def test_pull_fluid_to_single_step(asph, cuda_lib, oracle32, default_params):
    """pull_fluid_to: 13 * unit(pull - x_i) on top of gravity (simulation.rs:991-1003)."""
    sc, pos, vel, mass = _corner_block(asph)
    params = _uniform_params(default_params, pull_fluid_to=[0.3, 0.2, 0.0])
    _one_step(asph, cuda_lib, oracle32, params, pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))


@pytest.mark.parametrize("variant", ["OnlyDensity", "np_after_div", "OnlyDensity+np_after_div"])
def test_hybrid_dfsph_variants_single_step(asph, cuda_lib, oracle32, default_params, variant):
    """hybrid_dfsph_density_source_term: OnlyDensity (simulation.rs:1712-1748) and the non-pressure forces applied after the
    divergence solve (simulation.rs:2502-2576)."""
    sc, pos, vel, mass = _corner_block(asph)
    kw = {}
    if "OnlyDensity" in variant:
        kw["hybrid_dfsph_density_source_term"] = "OnlyDensity"
    if "np_after_div" in variant:
        kw["hybrid_dfsph_non_pressure_accel_before_divergence_free"] = False
    _one_step(asph, cuda_lib, oracle32, _uniform_params(default_params, **kw), pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))


@pytest.mark.parametrize("penalty", ["None", "Linear", "Quadratic2"])
@pytest.mark.parametrize("kind", ["AnalyticOverestimate", "AnalyticUnderestimate"])
def test_boundary_penalty_terms_single_step(asph, cuda_lib, oracle32, default_params, penalty, kind):
    """boundary_penalty_term (boundary_winchenbach2020.rs:87-118): the block overlaps the wall slightly (d < 0 for its outer
    particles), so every branch of the penalty and its derivative is taken; lambda sums compared as well."""
    sc, pos, vel, mass = _corner_block(asph, pos=(-1.003, -1.003))
    params = _uniform_params(default_params, boundary_penalty_term=penalty, init_boundary_handler=kind)
    b = asph.scene_boundary(sc, kind)
    g, o = _pair(asph, cuda_lib, oracle32, params, pos, vel, mass, b)
    g.single_step_without_adaptivity(); o.single_step_without_adaptivity()
    assert np.allclose(g.get_field("lambda_sum"), o.get_field("lambda_sum"), rtol=1e-6, atol=1e-7)
    assert np.allclose(g.get_field("lambda_grad"), o.get_field("lambda_grad"), rtol=1e-6, atol=1e-5 * max(1.0, float(np.abs(o.get_field("lambda_grad")).max())))
    g.close(); o.close()
    _one_step(asph, cuda_lib, oracle32, params, pos, vel, mass, b)


# ---- sizing_function ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sizing", ["Mass", "Radius2"])
@pytest.mark.parametrize("phase", ["share+merge", "share+split"])
def test_sizing_functions_resampling_parity(asph, cuda_lib, oracle32, default_params, split_patterns, sizing, phase):
    """target_mass with the Mass / Radius2 sizing functions (simulation.rs:213-237) decides classes, eligibility and split
    counts: single_step_adaptivity on identical inputs ends in the BIT-identical particle set."""
    pos, vel, mass, level, params = _adaptive_case(asph, default_params, 3)
    params = params.replace(sizing_function=sizing, sharing=True, merging="merge" in phase, splitting="split" in phase)
    b = asph.scene_boundary(_scene(asph, "default-scene.yaml"), "AnalyticOverestimate")
    g, o = _pair(asph, cuda_lib, oracle32, params, pos, vel, mass, b, split_patterns)
    _hooks(g, o)
    g.build_neighbors(np.float32(2.0)); o.build_neighbors(np.float32(2.0))
    lp = level.ctypes.data_as(C.POINTER(C.c_float))
    assert g.lib.asph_set_level(g._h, lp, len(level)) == 0
    assert o.lib.oracle_set_level(o._h, lp, len(level)) == 0
    step_number = 2 if "merge" in phase else 3
    g.lib.asph_set_step_number(g._h, step_number); o.lib.oracle_set_step_number(o._h, step_number)
    g.single_step_adaptivity(dt=0.002); o.single_step_adaptivity(dt=0.002)
    gi, oi = g.step_info(), o.step_info()
    assert (gi["n_shared"], gi["n_merged"], gi["n_split_parents"]) == (oi["n_shared"], oi["n_merged"], oi["n_split_parents"]), (gi, oi)
    assert oi["n_shared"] + oi["n_merged"] + oi["n_split_parents"] > 0
    assert g.num_fluid_particles() == o.num_fluid_particles()
    for f in ("mass", "position", "velocity"):
        assert np.array_equal(g.get_field(f), o.get_field(f)), f
    g.close(); o.close()


# ---- support_length_estimation != FromMass -----------------------------------------------------------------------------
H_MODES = ["FromDistribution", "FromDistributionClamped1", "FromDistributionClamped2", "FromDistribution2"]


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


@pytest.mark.parametrize("mode", ["FromDistributionClamped1", "FromDistribution"])
def test_support_length_from_distribution_with_resampling(asph, cuda_lib, oracle32, default_params, split_patterns, mode):
    """C1 with share / merge / split, 12 full steps: h2_next follows the particles through the resampling kernels."""
    _steps_with_resampling(asph, cuda_lib, oracle32, default_params.replace(support_length_estimation=mode), split_patterns, 12,
                           RESAMPLING_KEYS, 1e-4)


# ---- pressure_solver_method: IISPH2 --------------------------------------------------------------------------------------
def test_iisph2_single_step_uniform(asph, cuda_lib, oracle32, default_params):
    sc, pos, vel, mass = _corner_block(asph)
    _one_step(asph, cuda_lib, oracle32, _uniform_params(default_params, pressure_solver_method="IISPH2"), pos, vel, mass,
              asph.scene_boundary(sc, "AnalyticOverestimate"))


def test_iisph2_default_scene_with_resampling(asph, cuda_lib, oracle32, default_params, split_patterns):
    """C1, 12 full steps: the Large particles of each resampling phase take the single-term omega in the next step."""
    _steps_with_resampling(asph, cuda_lib, oracle32, default_params.replace(pressure_solver_method="IISPH2"), split_patterns, 12,
                           RESAMPLING_KEYS + ("density_sweeps",), 1e-5)


# ---- BASELINE configs[3]: ratio-stress-test geometry, radius ratio 16:1, IISPH --------------------------------------------
def _ratio_params(default_params):
    # media/ratio-stress-test.yaml:6-16, with the solver of BASELINE configs[3]
    return default_params.replace(merging=False, sharing=False, splitting=False, support_length_estimation="FromMass", max_iters=200,
                                  hybrid_dfsph_max_avg_density_error=0.001, hybrid_dfsph_max_avg_divergence_error=0.0001,
                                  hybrid_dfsph_factor=1000000, cfl_factor=0.3, max_dt=0.003, pressure_solver_method="IISPH")


def _ratio_scene(asph, fine, ratio, coarse_x):
    return asph.SceneConfig({"boundary": {"type": "box", "width": 2, "height": 2},
                             "blocks": [{"spacing": fine * ratio, "volume_fill_ratio": 0.93, "pos": [coarse_x, -0.5], "size": [0.55, 1.4], "velocity": [0, 0]},
                                        {"spacing": fine, "volume_fill_ratio": 0.93, "pos": [-0.95, -0.5], "size": [0.55, 1.4], "velocity": [0, 0]}]})


def test_ratio_stress_geometry_iisph(asph, cuda_lib, oracle32, default_params):
    """The scene of configs[3] (ratio-stress-test-scene.yaml: a coarse and a fine block side by side in the tank) at radius
    ratio 16:1 with fine spacing 0.008 (12 k particles), level estimation on, IISPH: 8 steps."""
    sc = _ratio_scene(asph, 0.008, 16.0, 0.4)
    params = _ratio_params(default_params)
    g = asph.init_fluid_sim(params, sc, None, lib=cuda_lib)
    o = asph.init_fluid_sim(params, sc, None, lib=oracle32)
    assert g.num_fluid_particles() > 11000
    for step in range(8):
        dg = g.single_step(); do = o.single_step()
        assert dg == do, step
        gi, oi = g.step_info(), o.step_info()
        assert (gi["density_sweeps"], gi["level_sweeps"]) == (oi["density_sweeps"], oi["level_sweeps"]), (step, gi, oi)
    assert np.array_equal(g.get_field("neighbor_count"), o.get_field("neighbor_count"))
    assert np.array_equal(g.get_field("flag_is_fluid_surface"), o.get_field("flag_is_fluid_surface"))
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= 1e-6
    assert _rel(g.get_field("density"), o.get_field("density"), 1.0) <= 2e-4
    g.close(); o.close()


def test_ratio_16_to_1_blocks_in_contact(asph, cuda_lib, oracle32, default_params):
    """The same two blocks pushed together, so that 16:1 pairs share neighbourhoods (five grid levels, far tables across
    levels, wide slices): neighbour sets bit-exact, then 3 IISPH steps."""
    sc = _ratio_scene(asph, 0.008, 16.0, -0.4 + 0.06)
    params = _ratio_params(default_params)
    g = asph.init_fluid_sim(params, sc, None, lib=cuda_lib)
    o = asph.init_fluid_sim(params, sc, None, lib=oracle32)
    g.build_neighbors(np.float32(2.0)); o.build_neighbors(np.float32(2.0))
    go, gi = g.neighbors_csr(); oo, oi = o.neighbors_csr()
    assert np.array_equal(go, oo) and np.array_equal(gi, oi)
    assert np.diff(oo.astype(np.int64)).max() > 100  # a coarse particle next to the fine block has hundreds of neighbours
    for step in range(3):
        dg = g.single_step(); do = o.single_step()
        assert dg == do, step
        assert g.step_info()["density_sweeps"] == o.step_info()["density_sweeps"], step
    assert np.array_equal(g.get_field("neighbor_count"), o.get_field("neighbor_count"))
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= 1e-6
    assert _rel(g.get_field("density"), o.get_field("density"), 1.0) <= 2e-4
    pmax = max(float(np.abs(o.get_field("pressure")).max()), 1e-6)
    assert _rel(g.get_field("pressure"), o.get_field("pressure"), pmax) <= 5e-4
    g.close(); o.close()


# ---- constrain_neighborhood_count (simulation.rs:2145-2177) ------------------------------------------------------------------
def _rings(asph, centers=((0.0, 0.0), (0.3, 0.1), (-0.4, 0.2)), n=21, seed=3):
    """Rings of 21 particles of diameter 1.3 h: every particle has the 20 others of its ring as neighbours (one more than
    optimal_neighbor_number + 5 = 19) and its second-farthest lies between h and 1.5 h away, so the reference's asserts on
    the new length (0 <= h_next < h) hold — for an ordinary particle distribution with a free surface they never do."""
    sc = asph.SceneConfig.dam_break(0.02)
    _, _, mass = asph.scene_particles(sc)
    m0 = np.float32(mass[0])
    h0 = 1.9 * np.sqrt(float(m0) / np.pi)  # rest_density 1
    rng = np.random.default_rng(seed)
    pts = []
    for c in centers:
        a = np.linspace(0, 2 * np.pi, n, endpoint=False) + rng.uniform(0, 1)
        r = 0.65 * h0 * (1 + rng.uniform(-0.03, 0.03, n))
        pts.append(np.stack([c[0] + r * np.cos(a), c[1] + r * np.sin(a)], 1))
    pos = np.concatenate(pts).astype(np.float32)
    return sc, pos, np.zeros_like(pos), np.full(len(pos), m0, np.float32)


@pytest.mark.parametrize("level", ["None", "EmptyAngle"])
def test_constrain_neighborhood_count_shrinks_h(asph, cuda_lib, oracle32, default_params, level):
    """The new smoothing lengths are bit-identical (the neighbour predicate of everything after depends on them), dt is
    taken from them (simulation.rs:2182-2191 comes after), and the step's fields agree as in every other mode.  The
    reference keeps its lists and lets the pairs outside the shrunken supports contribute zeros; the CUDA path rebuilds
    the lists, so neighbor_count is not compared."""
    sc, pos, vel, mass = _rings(asph)
    params = default_params.replace(constrain_neighborhood_count=True, sharing=False, merging=False, splitting=False,
                                    level_estimation_method=level)
    g, o = _pair(asph, cuda_lib, oracle32, params, pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))
    dg = g.single_step_without_adaptivity(); do = o.single_step_without_adaptivity()
    assert dg == do
    hg, ho = g.get_field("h"), o.get_field("h")
    assert np.array_equal(hg, ho)
    h0 = 1.9 * np.sqrt(float(mass[0]) / np.pi)
    assert ho.max() < 0.7 * h0  # every particle was over the target and got a shorter length
    gi, oi = g.step_info(), o.step_info()
    assert (gi["div_sweeps"], gi["density_sweeps"]) == (oi["div_sweeps"], oi["density_sweeps"]), (gi, oi)
    _compare_step_fields(g, o, 2e-4, [("density", 1.0), ("aii", None), ("ppe_source_term", None)])
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= 1e-6
    if level != "None":
        assert np.array_equal(g.get_field("flag_is_fluid_surface"), o.get_field("flag_is_fluid_surface"))
        assert _rel(g.get_field("level"), o.get_field("level"), 1.0) <= 1e-5
    g.close(); o.close()


def test_constrain_neighborhood_count_assert_like_the_reference(asph, cuda_lib, oracle32, default_params):
    """On a particle distribution with a free surface the reference's assert!(*p_h_next < h) fires; so do the oracle and the
    CUDA path, with the same error code."""
    sc, pos, vel, mass = _mixed_cloud(asph)
    params = default_params.replace(constrain_neighborhood_count=True, sharing=False, merging=False, splitting=False)
    g, o = _pair(asph, cuda_lib, oracle32, params, pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))
    codes = []
    for sim in (g, o):
        with pytest.raises(asph.AsphError) as e:
            sim.single_step_without_adaptivity()
        assert "constrain_neighborhood_count" in str(e.value)
        codes.append(e.value.code)
    assert codes[0] == codes[1]
    g.close(); o.close()


# ---- level_estimation_after_advection (simulation.rs:2678-2707) and CenterDiff (simulation.rs:631-695) -------------------------
@pytest.mark.parametrize("method", ["EmptyAngle", "CenterDiff"])
def test_level_estimation_after_advection_single_step(asph, cuda_lib, oracle32, default_params, method):
    """One step of the mixed-size cloud: the level set is estimated on lists rebuilt (extended range) at the advected
    positions, smoothed with the step's densities; the per-step fields follow the particles through the second sort."""
    sc, pos, vel, mass = _mixed_cloud(asph)
    params = default_params.replace(level_estimation_after_advection=True, level_estimation_method=method, sharing=False, merging=False,
                                    splitting=False)
    g, o = _pair(asph, cuda_lib, oracle32, params, pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))
    dg = g.single_step_without_adaptivity(); do = o.single_step_without_adaptivity()
    assert dg == do
    gi, oi = g.step_info(), o.step_info()
    assert (gi["div_sweeps"], gi["density_sweeps"], gi["level_sweeps"]) == (oi["div_sweeps"], oi["density_sweeps"], oi["level_sweeps"]), (gi, oi)
    pmax = max(float(np.abs(o.get_field("pressure")).max()), 1e-6)
    _compare_step_fields(g, o, 2e-4, [("density", 1.0), ("aii", None), ("ppe_source_term", None), ("pressure", pmax)])
    assert _rel(g.get_field("position"), o.get_field("position"), 2.0) <= 1e-6
    assert np.array_equal(g.get_field("flag_is_fluid_surface"), o.get_field("flag_is_fluid_surface"))
    lg, lo = g.get_field("level"), o.get_field("level")
    assert np.array_equal(lg > 0, lo > 0)  # the same particles are FluidInterior
    assert _rel(lg, lo, 1.0) <= 1e-5
    if method == "CenterDiff":
        assert (lo[lo <= 0] < 0).any() and o.get_field("flag_is_fluid_surface").sum() > 0
    g.close(); o.close()


@pytest.mark.parametrize("method", ["EmptyAngle", "CenterDiff"])
def test_level_estimation_after_advection_default_scene_with_resampling(asph, cuda_lib, oracle32, default_params, split_patterns, method):
    """C1, 12 full steps: the resampling phase runs on the lists of the second neighbour pass.  Identical particle counts,
    resampling statistics, density and level sweep counts every step.  (The divergence solve of the free-falling blocks is
    degenerate — the divergence is rounding noise, and whether ANY particle ends a sweep with positive pressure decides
    between one sweep and three: not compared here.)"""
    params = default_params.replace(level_estimation_after_advection=True, level_estimation_method=method)
    _steps_with_resampling(asph, cuda_lib, oracle32, params, split_patterns, 12,
                           RESAMPLING_KEYS + ("density_sweeps", "level_sweeps"), 1e-5)


def test_center_diff_before_advection_is_refused_like_the_reference(asph, cuda_lib, oracle32, default_params):
    """simulation.rs:2021: CenterDiff needs densities, which the level estimation at the top of the step does not have."""
    sc, pos, vel, mass = _corner_block(asph)
    params = default_params.replace(level_estimation_method="CenterDiff", sharing=False, merging=False, splitting=False)
    g, o = _pair(asph, cuda_lib, oracle32, params, pos, vel, mass, asph.scene_boundary(sc, "AnalyticOverestimate"))
    for sim in (g, o):
        with pytest.raises(asph.AsphError):
            sim.single_step_without_adaptivity()
    g.close(); o.close()


# ---- the two stage-fill variants of the single-GPU sweep kernels ---------------------------------------------------------
def test_bulk_copy_kernels_equal_per_thread_copy_kernels_bit_for_bit(asph, cuda_lib, default_params, monkeypatch):
    """k_sweep_bulk (cp.async.bulk + mbarrier, the default) and k_sweep (LDGSTS, ASPH_BULK=0) run the same arithmetic in
    the same order: 10 steps of a 78 k-particle uniform dam break (about 300 interior tiles) end in the bit-identical state."""
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


def test_bulk_copy_kernels_adaptive(asph, cuda_lib, default_params, split_patterns, monkeypatch):
    """The {h, m}-window instantiations: 8 steps of the adaptive mid-size dam break, with and without the bulk copies.
    With several size levels a step is not bit-reproducible from run to run (which columns get the slots of a crowded far
    table is decided by the arrival order of an atomic, and with it the order of a few fp32 sums), so the two stage fills
    are compared the way the step is compared with the oracle: same particle counts and resampling statistics every step,
    positions within 1e-6 of the domain size."""
    spacing = 0.004
    sc = asph.SceneConfig.dam_break(spacing)
    r_f = float(np.sqrt(0.93 / np.pi) * spacing)
    params = default_params.replace(particle_radius_fine=r_f, particle_radius_base=4.0 * r_f, maximum_surface_distance=0.2)
    out = []
    for flag in ("1", "0"):
        monkeypatch.setenv("ASPH_BULK", flag)
        g = asph.init_fluid_sim(params, sc, split_patterns, lib=cuda_lib)
        stats = []
        for _ in range(8):
            g.single_step()
            i = g.step_info()
            stats.append(tuple(i[k] for k in RESAMPLING_KEYS + ("div_sweeps", "density_sweeps", "level_sweeps")))
        out.append((stats, g.get_field("position"), g.get_field("mass")))
        g.close()
    assert out[0][0] == out[1][0]
    assert _rel(out[0][1], out[1][1], 2.0) <= 1e-6
    assert _rel(out[0][2], out[1][2]) <= 1e-6


# ---- north star: 1e-5 relative after 100 steps ------------------------------------------------------------------------------
def test_hundred_steps_c1_free_fall_with_resampling(asph, cuda_lib, oracle32, oracle64, default_params, split_patterns):
    """100 full steps (level set, share / merge / split: 1 035 -> 3 978 particles) of C1 with max_dt = 0.002, i.e. up to
    t = 0.2 s, before the two blocks land — the part of the trajectory that is not chaotic (after the impact the fp32 and fp64
    builds of the same oracle are 6e-2 of the domain apart at step 100).  Particle counts and resampling
    statistics identical every step; |x_gpu - x_oracle32| / L <= 1e-5 with the fp32-vs-fp64 oracle distance as the floor."""
    sc = _scene(asph, "default-scene.yaml")
    params = default_params.replace(max_dt=0.002)
    g = asph.init_fluid_sim(params, sc, split_patterns, lib=cuda_lib)
    o = asph.init_fluid_sim(params, sc, split_patterns, lib=oracle32)
    d = asph.init_fluid_sim(params, sc, split_patterns, lib=oracle64)
    for step in range(100):
        g.single_step(); o.single_step(); d.single_step()
        gi, oi = g.step_info(), o.step_info()
        for k in RESAMPLING_KEYS + ("div_sweeps", "density_sweeps", "level_sweeps"):
            assert gi[k] == oi[k], (step, k, gi, oi)
    assert abs(g.time - 0.2) < 1e-3
    err = _rel(g.get_field("position"), o.get_field("position"), 2.0)
    floor = _rel(o.get_field("position"), d.get_field("position"), 2.0) if o.num_fluid_particles() == d.num_fluid_particles() else float("nan")
    verr = _rel(g.get_field("velocity"), o.get_field("velocity"), max(float(np.abs(o.get_field("velocity")).max()), 1e-3))
    print(f"C1 100 steps to t = {g.time:.3f}: N = {g.num_fluid_particles()}, |gpu - oracle32| / L = {err:.3e} (velocity {verr:.3e}); "
          f"|oracle32 - oracle64| / L = {floor:.3e}")
    assert err <= 1e-5
    assert verr <= 1e-4
    g.close(); o.close(); d.close()
