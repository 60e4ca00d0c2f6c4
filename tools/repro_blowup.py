"""Debug helper: run C2 to a given step, save the state; or load a state and keep stepping (e.g. under compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import asph_b200 as A
from bench import uniform_params, dam_break, SPACING_C2
mode, path, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
params = uniform_params(A); scene = dam_break(A, SPACING_C2)
b = A.scene_boundary(scene, "AnalyticOverestimate")
if mode == "save":
    pos, vel, mass = A.scene_particles(scene)
    sim = A.FluidSimulation(params, pos, vel, mass, b, counters_enabled=True)
    for k in range(steps):
        sim.single_step()
    np.savez(path, pos=sim.get_field("position"), vel=sim.get_field("velocity"), mass=sim.get_field("mass"))
    print("saved at step", steps, "t", sim.time)
else:
    d = np.load(path)
    sim = A.FluidSimulation(params, d["pos"], d["vel"], d["mass"], b, counters_enabled=True)
    for k in range(steps):
        try:
            dt = sim.single_step(); i = sim.step_info()
            v = sim.get_field("velocity"); x = sim.get_field("position")
            print("step %d dt=%.2e div=%d den=%d vmax=%.3g xmax=%.3g" % (k, dt, i["div_sweeps"], i["density_sweeps"], np.abs(v).max(), np.abs(x).max()), flush=True)
        except Exception as e:
            print("FAILED at", k, e, flush=True); break
