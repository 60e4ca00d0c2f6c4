"""Exploration: when does the pressure solver engage in C2, and what does a step cost then?"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import asph_b200 as A
from bench import uniform_params, dam_break, SPACING_C2
spacing = float(sys.argv[1]) if len(sys.argv) > 1 else SPACING_C2
max_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
after = int(sys.argv[3]) if len(sys.argv) > 3 else 60
lib = A.load_library(); params = uniform_params(A); scene = dam_break(A, spacing)
pos, vel, mass = A.scene_particles(scene)
sim = A.FluidSimulation(params, pos, vel, mass, A.scene_boundary(scene, "AnalyticOverestimate"), counters_enabled=True, lib=lib)
n = len(mass); t0 = time.perf_counter(); engaged = None; rows = []
for k in range(max_steps):
    c0 = sim.counters()["simulation-step"][0]
    try:
        dt = sim.single_step(); i = sim.step_info()
    except Exception as e:
        print("FAILED at step", k, e); break
    ms = sim.counters()["simulation-step"][0] - c0
    rows.append((k, sim.time, dt, i["div_sweeps"], i["density_sweeps"], ms))
    if engaged is None and (i["div_sweeps"] > 3 or i["density_sweeps"] > 3):
        engaged = k
    if engaged is not None and k > engaged + after:
        break
print("particles", n, "engaged at step", engaged, "wall", time.perf_counter() - t0)
for r in rows[:3] + rows[max(0, (engaged or 0) - 3):]:
    print("step %d t=%.4f dt=%.2e div=%d den=%d ms=%.3f" % r)
