"""The merge transient of the bench workload, a few times over: does every run go through, with the same statistics?"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asph_b200 as A
import bench

def main():
    spacing = float(sys.argv[1]) if len(sys.argv) > 1 else bench.SPACING_16M
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    base = bench.adaptive_params(A, spacing)
    scene = bench.dam_break(A, spacing)
    base = A.init_simulation_params(base, scene)
    split = A.load_split_patterns_from_file()
    for rep in range(reps):
        sim = A.init_fluid_sim(base, scene, split, counters_enabled=True)
        rows = []
        try:
            for s in range(steps):
                sim.single_step(bench.ramped(base, spacing, s * 8))  # a fast ramp: the big merges within a few steps
                i = sim.step_info()
                lv = sim.get_field("level")
                rows.append((int(i["n_particles_end"]), int(i["level_sweeps"]), int(i["n_shared"]), int(i["n_merged"]), int(i["n_split_parents"]), sim.adapt_rounds(),
                             int(np.isnan(lv).sum()), float(np.nanmin(lv))))
        except A.AsphError as e:
            print(json.dumps({"rep": rep, "failed_at": len(rows), "error": str(e), "rows": rows[-3:]}), flush=True)
            sim.close()
            continue
        print(json.dumps({"rep": rep, "rows": rows}), flush=True)
        sim.close()

main()
