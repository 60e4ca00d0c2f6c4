"""Writes the input fixtures of this repository from the reference's shipped input files (run in the build container,
where /root/reference is mounted; the GPU box only sees the committed results).

The step loop is only comparable with the reference on the reference's own inputs: its default parameter sets, its
demo scenes and — for the split phase — the 58 precomputed split patterns its offline optimiser produced
(adaptivity/splitting.rs:84-120, 463-548; the optimiser is out of scope).  These are parameter VALUES, not code; the
fixtures carry them in this repository's own canonical layout, the way golden vectors are carried:

  configs/default-config.yaml, default-config-web.yaml    parameters grouped by the stage of the step that reads them
  configs/*-scene*.yaml                                    scenes, one block per line
  adaptive-sph_b200/data/split-patterns.yaml               one pattern per line (flow style), fp32 shortest round-trip digits

Both the Python loaders and the native host (yaml_lite.hpp) read these as well as the reference's original files;
tests/test_input_fixtures.py checks (in the build container) that every value equals the reference's.
"""
import os
import sys

import numpy as np
import yaml

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GROUPS = [
    ("fluid and time step (simulation.rs:2182-2191)", ["rest_density", "gravity", "cfl_factor", "max_dt", "h"]),
    ("neighbour search and smoothing length (neighborhood_search.rs, simulation.rs:1865-1971, 2145-2177)",
     ["neighborhood_search_algorithm", "support_length_estimation", "constrain_neighborhood_count", "check_neighborhood"]),
    ("boundary handler (boundary_winchenbach2020.rs, sdf/)", ["init_boundary_handler", "boundary_penalty_term", "sdf_gradient_eps"]),
    ("non-pressure forces (simulation.rs:931-1005)", ["viscosity_type", "viscosity"]),
    ("pressure solvers (simulation.rs:1207-1516, 2262-2670)",
     ["pressure_solver_method", "use_iisph", "operator_discretization", "jacobi_omega", "max_iters", "check_aii", "iisph_max_avg_density_error",
      "hybrid_dfsph_factor", "hybrid_dfsph_max_avg_density_error", "hybrid_dfsph_max_avg_divergence_error", "hybrid_dfsph_density_source_term",
      "hybrid_dfsph_non_pressure_accel_before_divergence_free", "eos_power", "eos_stiffness"]),
    ("level set (simulation.rs:539-927)",
     ["level_estimation_method", "level_estimation_range", "use_extended_range_for_level_estimation", "level_estimation_after_advection",
      "maximum_range", "boundary_is_fluid_surface"]),
    ("sizing function (simulation.rs:213-237)", ["sizing_function", "particle_radius_fine", "particle_radius_base", "maximum_surface_distance"]),
    ("resampling: share / merge / split (adaptivity/)",
     ["sharing", "merging", "splitting", "minimum_share_partners", "minimum_merge_partners", "max_mass_transfer_sharing", "max_mass_transfer_merging",
      "max_share_distance", "max_merge_distance", "allow_share_with_optimal_particle", "allow_share_with_too_small_particle",
      "allow_merge_with_optimal_particle", "allow_merge_on_size_difference", "fail_on_missing_split_pattern"]),
]


def scalar(v):
    if isinstance(v, bool):
        return "true" if v else "false"
    if v is None:
        return "null"
    if isinstance(v, float):
        return np.format_float_positional(v, unique=True, trim="0")  # "0.00001", not "1e-05" (YAML 1.1 wants a dot in a float)
    return str(v)


def write_params(src, dst, title):
    with open(src) as f:
        m = yaml.safe_load(f)
    left = dict(m)
    out = [f"# {title}", "# SimulationParams (simulation_parameters.rs:25-108): the values of the reference's shipped parameter set, grouped by",
           "# the stage of the step that reads them.  Written by tools/make_input_fixtures.py.", ""]
    for head, keys in GROUPS:
        out.append(f"# ---- {head}")
        for k in keys:
            out.append(f"{k}: {scalar(left.pop(k))}")
        out.append("")
    if left:
        out.append("# ---- other")
        for k in sorted(left):
            out.append(f"{k}: {scalar(left.pop(k))}")
    with open(dst, "w") as f:
        f.write("\n".join(out).rstrip() + "\n")


def write_scene(src, dst, title):
    with open(src) as f:
        m = yaml.safe_load(f)
    b = m["boundary"]
    out = [f"# {title}", "# SceneConfig (simulation.rs:3052-3072): tank and fluid blocks of the reference's scene.  Written by tools/make_input_fixtures.py.",
           f"boundary: {{type: {b['type']}, width: {scalar(b['width'])}, height: {scalar(b['height'])}}}", "blocks:"]
    for blk in m["blocks"]:
        vec = lambda v: "[" + ", ".join(scalar(x) for x in v) + "]"
        out.append(f"  - {{spacing: {scalar(blk['spacing'])}, volume_fill_ratio: {scalar(blk['volume_fill_ratio'])}, pos: {vec(blk['pos'])}, "
                   f"size: {vec(blk['size'])}, velocity: {vec(blk['velocity'])}}}")
    with open(dst, "w") as f:
        f.write("\n".join(out) + "\n")


def write_split_patterns(src, dst):
    with open(src) as f:
        pats = yaml.load(f, Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))
    out = ["# Split patterns: entry k is the 1 -> (k + 2) pattern (adaptivity/splitting.rs:84-120): child masses and smoothing lengths",
           "# relative to a unit parent, child positions in units of the parent's radius.  The 58 patterns are the output of the",
           "# reference's offline optimiser (splitting.rs:463-548, out of scope here); only pos_s is read at run time.",
           "# One pattern per line.  Written by tools/make_input_fixtures.py."]
    for p in pats:
        num = lambda v: np.format_float_positional(float(v), unique=True, trim="0")
        ms = "[" + ", ".join(num(v) for v in p["mass_s"]) + "]"
        hs = "[" + ", ".join(num(v) for v in p["h_s"]) + "]"
        ps = "[" + ", ".join("[" + num(x) + ", " + num(y) + "]" for x, y in p["pos_s"]) + "]"
        out.append(f"- {{mass_s: {ms}, pos_s: {ps}, h_s: {hs}}}")
    with open(dst, "w") as f:
        f.write("\n".join(out) + "\n")


def main():
    if not os.path.isdir(REF):
        sys.exit("the reference tree is not mounted here")
    cfg = os.path.join(ROOT, "configs")
    write_params(os.path.join(REF, "default-config.yaml"), os.path.join(cfg, "default-config.yaml"), "Default parameter set of the desktop build")
    write_params(os.path.join(REF, "default-config-web.yaml"), os.path.join(cfg, "default-config-web.yaml"), "Default parameter set of the web demo")
    for name, src, title in (("default-scene.yaml", "default-scene.yaml", "Two blocks of different resolution (C1)"),
                             ("default-scene-web.yaml", "default-scene-web.yaml", "Dam break of the web demo (geometry of the BASELINE dam-break configs)"),
                             ("motivation-scene2.yaml", "media/motivation-scene2.yaml", "Wide dam break of the motivation video"),
                             ("ratio-stress-test-scene.yaml", "media/ratio-stress-test-scene.yaml", "Two blocks with a 50 : 1 spacing ratio")):
        write_scene(os.path.join(REF, src), os.path.join(cfg, name), title)
    write_split_patterns(os.path.join(REF, "split-patterns.yaml"), os.path.join(ROOT, "adaptive-sph_b200", "data", "split-patterns.yaml"))


if __name__ == "__main__":
    main()
