"""Per-step diagnostics of the adaptive dam break (the recipe of bench.py without the ramp): particle count, level-set sweeps, greedy
rounds of the partner searches, resampling statistics and the device time of every PerformanceCounters label, one JSON
line per step.

  python tools/diag_adaptive.py --spacing 2.806e-4 --steps 140 [--every 1]
"""
import argparse
import ctypes as C
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import asph_b200 as A
from bench import dam_break


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spacing", type=float, default=5.612e-4)
    ap.add_argument("--steps", type=int, default=140)
    ap.add_argument("--ratio", type=float, default=4.0)
    ap.add_argument("--every", type=int, default=1)
    args = ap.parse_args()
    lib = A.load_library()
    lib.asph_adapt_rounds.argtypes = [C.c_void_p]
    lib.asph_adapt_rounds.restype = C.c_uint64
    r_f = math.sqrt(0.93 / math.pi) * args.spacing
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    params = A.SimulationParams.from_yaml(os.path.join(root, "configs", "default-config.yaml"))
    params = params.replace(particle_radius_fine=r_f, particle_radius_base=args.ratio * r_f, maximum_surface_distance=0.2)
    scene = dam_break(A, args.spacing)
    params = A.init_simulation_params(params, scene)
    sim = A.init_fluid_sim(params, scene, A.load_split_patterns_from_file(), counters_enabled=True, lib=lib)
    prev = {k: v[0] for k, v in sim.counters().items()}
    for s in range(args.steps):
        try:
            sim.single_step(params)
        except A.AsphError as e:
            print(json.dumps({"step": s, "failed": str(e)}))
            return 1
        cur = {k: v[0] for k, v in sim.counters().items()}
        i = sim.step_info()
        if s % args.every == 0:
            print(json.dumps({
                "step": s, "n": int(i["n_particles_begin"]), "n_end": int(i["n_particles_end"]), "t": round(sim.time, 5),
                "sweeps": [int(i["div_sweeps"]), int(i["density_sweeps"])], "level_sweeps": int(i["level_sweeps"]),
                "rounds": int(lib.asph_adapt_rounds(sim._h)), "shared": int(i["n_shared"]), "merged": int(i["n_merged"]),
                "split": int(i["n_split_parents"]), "ms": {k: round(cur[k] - prev[k], 3) for k in cur}}), flush=True)
        prev = cur
    sim.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
