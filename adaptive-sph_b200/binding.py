"""ctypes binding of include/asph.h and the host-side mirror of the reference's step API.

`FluidSimulation` mirrors the reference type of the same name (src/simulation/simulation.rs:471-537, 1973-2796):
`init_fluid_sim` builds it from params + scene + split patterns (simulation.rs:3074), `single_step`,
`single_step_without_adaptivity`, `single_step_adaptivity` step it, `particles` reads the SoA back in
reference particle order, `write_statistics` formats the counters (simulation.rs:3279).

The product library is adaptive-sph_b200/csrc/libasph_b200.so (hand-written sm_100a CUDA).  There is NO CPU
fallback: if the library is missing or no GPU is usable the constructor raises.  `load_library(path)` can bind
any other library exporting the same ABI; tests use that to drive the CPU oracle through this same class.
"""
import ctypes as C
import os

import numpy as np

from .params import AsphParams, SimulationParams
from .scene import AsphBoundary, SceneConfig, init_simulation_params, scene_boundary, scene_particles
from .split_patterns import AsphSplitPatterns, SplitPatterns, load_split_patterns_from_file

HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(HERE, "csrc", "libasph_b200.so")

ASPH_PC_LABELS = ["simulation-step", "neighborhood", "level-estimation", "div-solver", "density-solver", "adaptivity"]
ERRORS = {0: "OK", 1: "INVALID", 2: "UNSUPPORTED", 3: "NONFINITE", 4: "NEG_AII", 5: "DENSITY", 6: "MASS_CONSERVATION",
          7: "NEIGHBOR_OVERFLOW", 8: "CUDA", 9: "NCCL", 10: "CAPACITY", 11: "NO_DEVICE"}

# field id -> (numpy dtype, components)
FIELDS = {
    "position": (0, np.float32, 2), "velocity": (1, np.float32, 2), "mass": (2, np.float32, 1),
    "h": (3, np.float32, 1), "density": (4, np.float32, 1), "pressure": (5, np.float32, 1),
    "aii": (6, np.float32, 1), "ppe_source_term": (7, np.float32, 1), "pressure_accel": (8, np.float32, 2),
    "level": (9, np.float32, 1), "particle_size_class": (10, np.uint8, 1), "neighbor_count": (11, np.uint32, 1),
    "flag_is_fluid_surface": (12, np.uint8, 1), "flag_insufficient_neighs": (13, np.uint8, 1),
    "lambda_sum": (14, np.float32, 1), "lambda_grad": (15, np.float32, 2), "merge_partner": (16, np.uint32, 1),
    "merge_counter": (17, np.uint16, 1), "density_error": (18, np.float32, 1), "constant_field": (19, np.float32, 1),
}
LEVEL_INTERIOR = 1.0


class AsphStepInfo(C.Structure):
    _fields_ = [
        ("dt", C.c_float),
        ("div_iterations", C.c_int32), ("density_iterations", C.c_int32),
        ("div_sweeps", C.c_int32), ("density_sweeps", C.c_int32), ("level_sweeps", C.c_int32),
        ("n_shared", C.c_int32), ("n_merged", C.c_int32), ("n_split_parents", C.c_int32),
        ("n_particles_begin", C.c_uint64), ("n_particles_end", C.c_uint64),
        ("last_avg_error_div", C.c_double), ("last_avg_error_density", C.c_double),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class AsphError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"asph error {code} ({ERRORS.get(code, '?')}): {msg}")
        self.code = code


_P = C.c_void_p
_FP = C.POINTER(C.c_float)


def _declare(lib):
    lib.asph_create.argtypes = [C.POINTER(AsphParams), _FP, _FP, _FP, C.c_uint64, C.POINTER(AsphBoundary),
                                C.POINTER(AsphSplitPatterns), C.c_int, C.c_uint64, C.POINTER(_P)]
    lib.asph_create.restype = C.c_int
    lib.asph_destroy.argtypes = [_P]
    lib.asph_destroy.restype = None
    lib.asph_set_state.argtypes = [_P, _FP, _FP, _FP, C.c_uint64]
    lib.asph_set_state.restype = C.c_int
    for name in ("asph_step", "asph_step_physics"):
        getattr(lib, name).argtypes = [_P, C.POINTER(AsphParams), _FP]
        getattr(lib, name).restype = C.c_int
    lib.asph_step_adaptivity.argtypes = [_P, C.POINTER(AsphParams), C.c_float]
    lib.asph_step_adaptivity.restype = C.c_int
    lib.asph_num_particles.argtypes = [_P]
    lib.asph_num_particles.restype = C.c_uint64
    lib.asph_time.argtypes = [_P]
    lib.asph_time.restype = C.c_double
    lib.asph_step_number.argtypes = [_P]
    lib.asph_step_number.restype = C.c_uint64
    lib.asph_get_field.argtypes = [_P, C.c_int, C.c_void_p, C.c_uint64]
    lib.asph_get_field.restype = C.c_int
    lib.asph_get_neighbors_csr.argtypes = [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_uint64,
                                           C.POINTER(C.c_uint64)]
    lib.asph_get_neighbors_csr.restype = C.c_int
    lib.asph_build_neighbors.argtypes = [_P, C.POINTER(AsphParams), C.c_float]
    lib.asph_build_neighbors.restype = C.c_int
    lib.asph_get_step_info.argtypes = [_P, C.POINTER(AsphStepInfo)]
    lib.asph_get_step_info.restype = C.c_int
    lib.asph_get_counters.argtypes = [_P, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    lib.asph_get_counters.restype = C.c_int
    lib.asph_last_error.argtypes = [_P]
    lib.asph_last_error.restype = C.c_char_p
    lib.asph_set_kernel_timing.argtypes = [_P, C.c_int]
    lib.asph_set_kernel_timing.restype = C.c_int
    lib.asph_get_kernel_timing.argtypes = [_P, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    lib.asph_get_kernel_timing.restype = C.c_int
    lib.asph_adapt_rounds.argtypes = [_P]
    lib.asph_adapt_rounds.restype = C.c_uint64
    lib.asph_debug_greedy_duplicates.argtypes = [_P]
    lib.asph_debug_greedy_duplicates.restype = C.c_uint64
    lib.asph_set_level.argtypes = [_P, _FP, C.c_uint64]
    lib.asph_set_level.restype = C.c_int
    lib.asph_set_step_number.argtypes = [_P, C.c_uint64]
    lib.asph_set_step_number.restype = None
    lib.asph_kernel_launches.argtypes = [_P]
    lib.asph_kernel_launches.restype = C.c_uint64
    lib.asph_backend_name.argtypes = []
    lib.asph_backend_name.restype = C.c_char_p
    lib.asph_kernel_w.argtypes = [C.c_float, C.c_float]
    lib.asph_kernel_w.restype = C.c_float
    lib.asph_kernel_grad.argtypes = [C.c_float, C.c_float, C.c_float, _FP, _FP]
    lib.asph_kernel_grad.restype = None
    for name in ("asph_lambda", "asph_dlambda"):
        getattr(lib, name).argtypes = [C.c_double]
        getattr(lib, name).restype = C.c_double
    for name in ("asph_lambda_lut", "asph_dlambda_lut"):
        getattr(lib, name).argtypes = [C.c_float]
        getattr(lib, name).restype = C.c_float
    lib.asph_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    lib.asph_comm_unique_id.restype = C.c_int
    lib.asph_create_distributed.argtypes = [C.POINTER(AsphParams), _FP, _FP, _FP, C.POINTER(C.c_uint32), C.c_uint64,
                                            C.c_uint64, C.POINTER(AsphBoundary), C.POINTER(AsphSplitPatterns), C.c_int,
                                            C.c_uint64, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int,
                                            C.POINTER(_P)]
    lib.asph_create_distributed.restype = C.c_int
    lib.asph_get_global_index.argtypes = [_P, C.POINTER(C.c_uint32), C.c_uint64]
    lib.asph_get_global_index.restype = C.c_int
    return lib


_LIBS = {}


def load_library(path=None):
    """Load a library exporting include/asph.h.  Default: the CUDA product library; raises if it is not built."""
    path = os.path.abspath(path or PRODUCT_LIB)
    if path not in _LIBS:
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback for the product path.")
        _LIBS[path] = _declare(C.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2)))
    return _LIBS[path]


def _fp(a):
    return a.ctypes.data_as(_FP)


class ParticleView:
    """Read-only SoA view in reference particle order (`ParticleVec`, simulation.rs:284-334)."""

    def __init__(self, sim):
        self._sim = sim

    def __getattr__(self, name):
        if name in FIELDS:
            return self._sim.get_field(name)
        raise AttributeError(name)


class FluidSimulation:
    def __init__(self, params, pos, vel, mass, boundary=None, split_patterns=None, counters_enabled=False,
                 capacity=0, lib=None, distributed=None):
        self.lib = lib if lib is not None else load_library()
        self.params = params
        self._split = split_patterns
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 2)
        vel = np.ascontiguousarray(vel, dtype=np.float32).reshape(-1, 2)
        mass = np.ascontiguousarray(mass, dtype=np.float32).reshape(-1)
        assert len(pos) == len(vel) == len(mass)  # simulation.rs:496-497
        self._boundary = boundary if boundary is not None else AsphBoundary()
        handle = _P()
        sp = C.byref(split_patterns.c) if split_patterns is not None else None
        if distributed is None:
            rc = self.lib.asph_create(C.byref(params.c), _fp(pos), _fp(vel), _fp(mass), len(mass),
                                      C.byref(self._boundary), sp, int(counters_enabled), int(capacity),
                                      C.byref(handle))
        else:
            gidx = np.ascontiguousarray(distributed["global_index"], dtype=np.uint32)
            nid = (C.c_uint8 * 128).from_buffer_copy(bytes(distributed["nccl_id"]))
            rc = self.lib.asph_create_distributed(
                C.byref(params.c), _fp(pos), _fp(vel), _fp(mass), gidx.ctypes.data_as(C.POINTER(C.c_uint32)),
                len(mass), int(distributed["n_global"]), C.byref(self._boundary), sp, int(counters_enabled),
                int(capacity), nid, int(distributed["rank"]), int(distributed["n_ranks"]),
                int(distributed["device"]), C.byref(handle))
        if rc != 0:
            raise AsphError(rc, "asph_create failed" + (
                " (no usable CUDA device; the product has no CPU path)" if rc == 11 else ""))
        self._h = handle
        self.particles = ParticleView(self)

    # ---- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.asph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.asph_last_error(self._h)
            raise AsphError(rc, msg.decode() if msg else "")

    # ---- stepping (simulation.rs:1973, 1980, 2732) --------------------------------------------------
    def single_step(self, params=None):
        p = params or self.params
        dt = C.c_float()
        self._check(self.lib.asph_step(self._h, C.byref(p.c), C.byref(dt)))
        return dt.value

    def single_step_without_adaptivity(self, params=None):
        p = params or self.params
        dt = C.c_float()
        self._check(self.lib.asph_step_physics(self._h, C.byref(p.c), C.byref(dt)))
        return dt.value

    def single_step_adaptivity(self, params=None, dt=0.0):
        p = params or self.params
        self._check(self.lib.asph_step_adaptivity(self._h, C.byref(p.c), C.c_float(dt)))

    # ---- state --------------------------------------------------------------------------------------
    def set_state(self, pos, vel, mass):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        vel = np.ascontiguousarray(vel, dtype=np.float32)
        mass = np.ascontiguousarray(mass, dtype=np.float32)
        self._check(self.lib.asph_set_state(self._h, _fp(pos), _fp(vel), _fp(mass), mass.size))

    def num_fluid_particles(self):
        return int(self.lib.asph_num_particles(self._h))

    @property
    def time(self):
        return float(self.lib.asph_time(self._h))

    @property
    def step_number(self):
        return int(self.lib.asph_step_number(self._h))

    def get_field(self, name, out=None):
        fid, dtype, comps = FIELDS[name]
        n = self.num_fluid_particles()
        if out is None:
            out = np.empty((n, comps) if comps > 1 else (n,), dtype=dtype)
        self._check(self.lib.asph_get_field(self._h, fid, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def neighbors_csr(self):
        nnz = C.c_uint64()
        self._check(self.lib.asph_get_neighbors_csr(self._h, None, None, 0, C.byref(nnz)))
        n = self.num_fluid_particles()
        offsets = np.empty(n + 1, dtype=np.uint64)
        idx = np.empty(max(1, nnz.value), dtype=np.uint32)
        self._check(self.lib.asph_get_neighbors_csr(self._h, offsets.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                    idx.ctypes.data_as(C.POINTER(C.c_uint32)), idx.size,
                                                    C.byref(nnz)))
        return offsets, idx[:nnz.value]

    def build_neighbors(self, range_factor, params=None):
        p = params or self.params
        self._check(self.lib.asph_build_neighbors(self._h, C.byref(p.c), C.c_float(range_factor)))

    def step_info(self):
        info = AsphStepInfo()
        self._check(self.lib.asph_get_step_info(self._h, C.byref(info)))
        return info.as_dict()

    def counters(self):
        ms = (C.c_double * 6)()
        calls = (C.c_uint64 * 6)()
        self._check(self.lib.asph_get_counters(self._h, ms, calls))
        return {ASPH_PC_LABELS[i]: (ms[i], calls[i]) for i in range(6)}

    def global_index(self):
        n = self.num_fluid_particles()
        out = np.empty(n, dtype=np.uint32)
        self._check(self.lib.asph_get_global_index(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32)), n))
        return out

    def set_kernel_timing(self, sample_every):
        self._check(self.lib.asph_set_kernel_timing(self._h, int(sample_every)))

    def kernel_timing(self):
        ms = (C.c_double * 6)()
        cnt = (C.c_uint64 * 6)()
        self._check(self.lib.asph_get_kernel_timing(self._h, ms, cnt))
        names = ["accel_sweep", "jacobi_sweep", "neighbors", "sort_grid", "level_propagate", "partner_search"]
        return {names[k]: (ms[k], cnt[k]) for k in range(6)}

    def greedy_duplicates(self):
        return int(self.lib.asph_debug_greedy_duplicates(self._h))

    def adapt_rounds(self):
        """Dependency rounds of the greedy partner searches in the last resampling phase (0 on the CPU oracle)."""
        return int(self.lib.asph_adapt_rounds(self._h))

    def kernel_launches(self):
        return int(self.lib.asph_kernel_launches(self._h))

    def backend(self):
        return self.lib.asph_backend_name().decode()


def init_fluid_sim(params, scene, split_patterns=None, counters_enabled=False, lib=None, capacity=0):
    """init_fluid_sim (simulation.rs:3074-3231)."""
    pos, vel, mass = scene_particles(scene)
    boundary = scene_boundary(scene, params["init_boundary_handler"])
    return FluidSimulation(params, pos, vel, mass, boundary, split_patterns, counters_enabled, capacity, lib)


class StatisticsRecorder:
    """ValueCounters + write_statistics (simulation.rs:137-157, 3279-3359)."""

    def __init__(self):
        self.values = {}

    def add(self, key, v):
        self.values.setdefault(key, []).append(float(v))

    def record_step(self, info):
        self.add("particle-count", info["n_particles_begin"])
        self.add("dt", info["dt"])
        if info["div_iterations"] > 0:
            self.add("div-iterations", info["div_iterations"])
        if info["density_iterations"] > 0:
            self.add("density-iterations", info["density_iterations"])

    def write_statistics(self, sim):
        counters = sim.counters()
        sim_ms, _ = counters["simulation-step"]
        avg = lambda k: (sum(self.values[k]) / len(self.values[k])) if self.values.get(k) else float("nan")
        s = []
        s.append("${:.2f}\\si{{\\second}}$ & {} & {:.02f} & {:.02f} & - \\\\".format(
            sim_ms / 1000.0, int(round(avg("particle-count"))) if self.values.get("particle-count") else 0,
            avg("div-iterations"), avg("density-iterations")))
        s.append("")
        s.append(f"simulation-time: {sim_ms}ms")
        s.append("")
        for label in sorted(counters):
            ms, calls = counters[label]
            if calls:
                s.append(f"{label}: avg:{ms / calls}ms")
        s.append("")
        for label in sorted(self.values):
            v = self.values[label]
            s.append(f"{label}: min:{min(v)} max:{max(v)} avg:{sum(v) / len(v)}")
        return "\n".join(s) + "\n"
