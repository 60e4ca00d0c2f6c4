"""SimulationParams: YAML <-> the C struct `asph_params` (include/asph.h).

Mirrors `SimulationParams` and its serde enums (reference src/simulation/simulation_parameters.rs:4-213):
field names are the YAML keys; every non-Option field is mandatory (serde would refuse the file), the three
Option fields (`pull_fluid_to`, `fill_stash_with`, `operator_discretization_for_diagonal`) may be absent.
The `-c overwrite.yaml` merge follows platform/desktop/main_loop.rs:113-126 (unknown key -> error).
"""
import ctypes as C

import yaml

ENUMS = {
    "viscosity_type": ["WCSPH", "ApproxLaplace", "XSPH"],
    "level_estimation_method": ["None", "CenterDiff", "EmptyAngle"],
    "neighborhood_search_algorithm": ["Grid", "RStar"],
    "init_boundary_handler": ["Particles", "AnalyticUnderestimate", "AnalyticOverestimate", "NoBoundary"],
    "support_length_estimation": ["FromDistribution", "FromDistributionClamped1", "FromDistributionClamped2",
                                  "FromDistribution2", "FromMass"],
    "pressure_solver_method": ["IISPH", "IISPH2", "HybridDFSPH", "OnlyDivergence"],
    "hybrid_dfsph_density_source_term": ["DensityAndDivergence", "OnlyDensity"],
    "boundary_penalty_term": ["None", "Linear", "Quadratic1", "Quadratic2"],
    "sizing_function": ["Radius2", "Radius", "Mass"],
    "operator_discretization": ["ConsistentSimpleGradient", "ConsistentSymmetricGradient", "Winchenbach2020"],
}
OPTIONAL_ENUMS = {
    # name -> (choices, value for None)
    "fill_stash_with": (["SurfaceDistanceFirstIteration", "SurfaceDistanceMiddle"], 0),  # stored 1-based
    "operator_discretization_for_diagonal": (ENUMS["operator_discretization"], -1),
}

_D, _I, _L = C.c_double, C.c_int32, C.c_int64


class AsphParams(C.Structure):
    """Binary layout of `asph_params` in include/asph.h (keep in the same order)."""
    _fields_ = [
        ("rest_density", _D), ("cfl_factor", _D), ("max_dt", _D), ("h", _D),
        ("use_iisph", _I),
        ("viscosity", _D),
        ("viscosity_type", _I),
        ("gravity", _D),
        ("check_aii", _I),
        ("level_estimation_method", _I),
        ("maximum_range", _D),
        ("jacobi_omega", _D),
        ("eos_stiffness", _D),
        ("eos_power", _I),
        ("neighborhood_search_algorithm", _I),
        ("init_boundary_handler", _I),
        ("support_length_estimation", _I),
        ("sdf_gradient_eps", _D),
        ("fail_on_missing_split_pattern", _I),
        ("has_pull_fluid_to", _I),
        ("pull_fluid_to", _D * 3),
        ("constrain_neighborhood_count", _I),
        ("particle_radius_fine", _D), ("particle_radius_base", _D), ("maximum_surface_distance", _D),
        ("minimum_share_partners", _I), ("minimum_merge_partners", _I),
        ("merging", _I), ("sharing", _I), ("splitting", _I),
        ("max_mass_transfer_sharing", _D), ("max_mass_transfer_merging", _D),
        ("max_share_distance", _D), ("max_merge_distance", _D),
        ("allow_merge_with_optimal_particle", _I), ("allow_share_with_optimal_particle", _I),
        ("allow_share_with_too_small_particle", _I), ("allow_merge_on_size_difference", _I),
        ("boundary_is_fluid_surface", _I), ("use_extended_range_for_level_estimation", _I),
        ("pressure_solver_method", _I),
        ("iisph_max_avg_density_error", _D), ("hybrid_dfsph_factor", _D),
        ("hybrid_dfsph_max_avg_density_error", _D), ("hybrid_dfsph_max_avg_divergence_error", _D),
        ("hybrid_dfsph_density_source_term", _I),
        ("hybrid_dfsph_non_pressure_accel_before_divergence_free", _I),
        ("check_neighborhood", _I),
        ("fill_stash_with", _I),
        ("boundary_penalty_term", _I),
        ("sizing_function", _I),
        ("level_estimation_after_advection", _I),
        ("level_estimation_range", _D),
        ("operator_discretization", _I),
        ("operator_discretization_for_diagonal", _I),
        ("max_iters", _L),
    ]


_SPECIAL = {"has_pull_fluid_to", "pull_fluid_to"} | set(OPTIONAL_ENUMS)
MANDATORY_KEYS = [n for n, _ in AsphParams._fields_ if n not in _SPECIAL]
OPTIONAL_KEYS = ["pull_fluid_to", "fill_stash_with", "operator_discretization_for_diagonal"]


class SimulationParams:
    """Python-side value object; `.c` is the C struct handed to the library BY VALUE each step."""

    def __init__(self, mapping):
        self.values = dict(mapping)
        unknown = set(self.values) - set(MANDATORY_KEYS) - set(OPTIONAL_KEYS)
        if unknown:
            raise KeyError(f"unknown SimulationParams field(s): {sorted(unknown)}")
        missing = [k for k in MANDATORY_KEYS if k not in self.values]
        if missing:
            raise KeyError(f"failed to unpack SimulationParams: missing field(s) {missing}")
        self.c = self._to_c()

    # ---- serde-compatible loading -------------------------------------------------------------
    @classmethod
    def from_yaml(cls, path, overwrite_path=None, overrides=None):
        with open(path) as f:
            mapping = yaml.safe_load(f)
        if overwrite_path is not None:
            with open(overwrite_path) as f:
                over = yaml.safe_load(f) or {}
            mapping = merge_overwrite(mapping, over)
        if overrides:
            mapping = merge_overwrite(mapping, overrides, allow_new_optional=True)
        return cls(mapping)

    def replace(self, **kw):
        m = dict(self.values)
        m.update(kw)
        return SimulationParams(m)

    def __getitem__(self, k):
        return self.values[k]

    def _to_c(self):
        c = AsphParams()
        for name, ctype in AsphParams._fields_:
            if name in _SPECIAL:
                continue
            v = self.values[name]
            if name in ENUMS:
                # serde unit variants are plain strings; YAML `None` parses to Python None for the variant "None"
                key = "None" if v is None else str(v)
                if key not in ENUMS[name]:
                    raise ValueError(f"{name}: unknown variant {v!r}, expected one of {ENUMS[name]}")
                setattr(c, name, ENUMS[name].index(key))
            elif ctype is _D:
                setattr(c, name, float(v))
            else:
                setattr(c, name, int(v))
        pull = self.values.get("pull_fluid_to")
        c.has_pull_fluid_to = 0 if pull is None else 1
        if pull is not None:
            for k in range(3):
                c.pull_fluid_to[k] = float(pull[k])
        for name, (choices, none_value) in OPTIONAL_ENUMS.items():
            v = self.values.get(name)
            if v is None:
                setattr(c, name, none_value)
            else:
                setattr(c, name, choices.index(str(v)) + (1 if name == "fill_stash_with" else 0))
        return c


def merge_overwrite(mapping, over, allow_new_optional=False):
    """`-c` semantics: every key of the overwrite file must already exist (main_loop.rs:119-124)."""
    out = dict(mapping)
    for k, v in over.items():
        if k not in out and not (allow_new_optional and k in OPTIONAL_KEYS):
            raise KeyError(f"not able to find attribute {k}")
        out[k] = v
    return out
