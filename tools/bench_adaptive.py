"""Adaptive dam break at scale (SURVEY.md §8d, C3 recipe): BASELINE configs[2] (spacing 5.612e-4, 3 999 129 particles) and the
north star's 16 M-particle case (spacing 2.806e-4) on ONE GPU — level set + share / merge / split every step.

  python tools/bench_adaptive.py [--spacing 5.612e-4] [--warmup 100] [--steps 100] [--ratio 4] [--lib PATH]

Recipe: the dam-break block of bench.py (gentle drop, see there) at the FINE resolution: particle_radius_fine = the radius of
an initial particle (sqrt(0.93 / pi) * spacing), particle_radius_base = ratio * fine, maximum_surface_distance 0.2, default
config otherwise (EmptyAngle level set, HybridDFSPH, resampling on).  The interior merges towards the base size during the
warm-up steps, so the particle count drifts down; value = sum of N_step over the timed steps / their wall time (every step
call returns synchronised).  Prints one JSON line with the per-label device times of the step (PerformanceCounters) beside it.
`--lib` binds any library exporting include/asph.h (default: the CUDA library).
"""
import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import asph_b200 as A
from bench import dam_break


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spacing", type=float, default=5.612e-4)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--ratio", type=float, default=4.0, help="particle_radius_base / particle_radius_fine")
    ap.add_argument("--lib", default=None)
    args = ap.parse_args()
    lib = A.load_library(args.lib) if args.lib else A.load_library()
    r_f = math.sqrt(0.93 / math.pi) * args.spacing
    params = A.SimulationParams.from_yaml(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "default-config.yaml"))
    params = params.replace(particle_radius_fine=r_f, particle_radius_base=args.ratio * r_f, maximum_surface_distance=0.2)
    scene = dam_break(A, args.spacing)
    params = A.init_simulation_params(params, scene)
    t0 = time.perf_counter()
    sim = A.init_fluid_sim(params, scene, A.load_split_patterns_from_file(), counters_enabled=True, lib=lib)
    n0 = sim.num_fluid_particles()
    setup_s = time.perf_counter() - t0
    failed = None
    counts, sweeps, lvl = [], [], []

    def run(k_steps, record):
        nonlocal failed
        for _ in range(k_steps):
            try:
                sim.single_step(params)
            except A.AsphError as e:
                failed = str(e)
                return
            if record:
                i = sim.step_info()
                counts.append(int(i["n_particles_begin"]))
                sweeps.append((int(i["div_sweeps"]), int(i["density_sweeps"])))
                lvl.append(int(i["level_sweeps"]))

    run(args.warmup, False)
    c0 = {k: v[0] for k, v in sim.counters().items()}
    t1 = time.perf_counter()
    if failed is None:
        run(args.steps, True)
    wall = time.perf_counter() - t1
    c1 = {k: v[0] for k, v in sim.counters().items()}
    done = len(counts)
    out = {
        "workload": f"adaptive dam break, spacing {args.spacing:g}, radius ratio {args.ratio:g}:1, level set + share / merge / split",
        "backend": sim.backend(), "particles_initial": int(n0), "particles_timed_first": counts[0] if counts else None,
        "particles_timed_last": counts[-1] if counts else None, "warmup": args.warmup, "steps": done,
        "value": (sum(counts) / wall) if done else None, "unit": "particle-steps/s", "ms_per_step": (1e3 * wall / done) if done else None,
        "sweeps_per_step": [float(np.mean([s[0] for s in sweeps])), float(np.mean([s[1] for s in sweeps]))] if done else None,
        "level_sweeps_per_step": float(np.mean(lvl)) if done else None,
        "device_ms_per_step": {k: (c1[k] - c0[k]) / done for k in c1} if done else None,
        "setup_s": setup_s, "failed": failed, "simulated_time": float(sim.time),
    }
    print(json.dumps(out))
    sim.close()
    return 0 if failed is None else 1


if __name__ == "__main__":
    sys.exit(main())
