#!/usr/bin/env python
"""Generates tests/golden/*.npz: small known-answer vectors of the step path.

The reference (Rust) cannot be built or imported in this image, and it ships no step-level fixtures,
so these vectors come from the fp64 build of the CPU oracle (oracle/liboracle_f64.so, the `double-precision` cargo
feature of the reference restated) and pin BOTH the fp32 oracle and the CUDA path against silent drift:
  c1_physics_5.npz     default-config + default-scene, resampling off, 5 physics steps: x, v, rho, level after step 5
  c1_resampling_12.npz default-config + default-scene + split-patterns, 12 full steps: particle count per step, x, m
  uniform_step.npz     dam-break at spacing 0.02, uniform-h recipe, 1 step: rho, a_ii, source, pressure, a^p, dt, sweeps
Run from the repo root:  python tools/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import asph_b200 as A  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def f64(sim, name, comps):
    """read a real field in double from the f64 oracle (oracle-only getter)"""
    lib = sim.lib
    lib.oracle_get_field_f64.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
    lib.oracle_get_field_f64.restype = C.c_int
    n = sim.num_fluid_particles()
    out = np.empty((n, comps) if comps > 1 else (n,), dtype=np.float64)
    assert lib.oracle_get_field_f64(sim._h, A.FIELDS[name][0], out.ctypes.data_as(C.c_void_p), out.nbytes) == 0
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    lib = A.load_library(os.path.join(ROOT, "oracle", "liboracle_f64.so"))
    params = A.SimulationParams.from_yaml(os.path.join(ROOT, "configs", "default-config.yaml"))
    scene = A.SceneConfig.from_yaml(os.path.join(ROOT, "configs", "default-scene.yaml"))
    split = A.load_split_patterns_from_file()

    p = params.replace(merging=False, sharing=False, splitting=False)
    s = A.init_fluid_sim(p, scene, None, lib=lib)
    dts, sweeps = [], []
    for _ in range(5):
        dts.append(s.single_step())
        i = s.step_info()
        sweeps.append((i["div_sweeps"], i["density_sweeps"], i["level_sweeps"]))
    np.savez_compressed(os.path.join(OUT, "c1_physics_5.npz"), position=f64(s, "position", 2), velocity=f64(s, "velocity", 2),
                        density=f64(s, "density", 1), level=f64(s, "level", 1), dt=np.array(dts), sweeps=np.array(sweeps))
    s.close()

    s = A.init_fluid_sim(params, scene, split, lib=lib)
    counts = []
    for _ in range(12):
        s.single_step()
        counts.append(s.num_fluid_particles())
    np.savez_compressed(os.path.join(OUT, "c1_resampling_12.npz"), counts=np.array(counts), position=f64(s, "position", 2),
                        mass=f64(s, "mass", 1))
    s.close()

    sc = A.SceneConfig.dam_break(0.02)
    pos, vel, mass = A.scene_particles(sc)
    rng = np.random.default_rng(1)
    vel = (rng.standard_normal(vel.shape) * 0.05).astype(np.float32)
    p = params.replace(merging=False, sharing=False, splitting=False, level_estimation_method="None")
    s = A.FluidSimulation(p, pos, vel, mass, A.scene_boundary(sc, "AnalyticOverestimate"), lib=lib)
    dt = s.single_step()
    i = s.step_info()
    np.savez_compressed(os.path.join(OUT, "uniform_step.npz"), vel0=vel, dt=dt, sweeps=np.array([i["div_sweeps"], i["density_sweeps"]]),
                        density=f64(s, "density", 1), aii=f64(s, "aii", 1), source=f64(s, "ppe_source_term", 1),
                        pressure=f64(s, "pressure", 1), pressure_accel=f64(s, "pressure_accel", 2), position=f64(s, "position", 2),
                        velocity=f64(s, "velocity", 2))
    s.close()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
