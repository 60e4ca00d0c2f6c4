"""Long-run stability probe: the resting dam break (fill 1.0, on the floor) at a given spacing on the GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import asph_b200 as A
from bench import uniform_params
spacing = float(sys.argv[1]); steps = int(sys.argv[2])
params = uniform_params(A)
scene = A.SceneConfig.dam_break(spacing, pos=(-0.95, -1 + 0.5 * spacing), size=(0.7, 1.8), fill=1.0)
pos, vel, mass = A.scene_particles(scene)
g = A.FluidSimulation(params, pos, vel, mass, A.scene_boundary(scene, "AnalyticOverestimate"))
for k in range(steps):
    try:
        dt = g.single_step()
    except Exception as e:
        print("FAILED at step", k, str(e)[:80], flush=True); break
    i = g.step_info()
    if k % 100 == 0 or i["density_sweeps"] > 100 or i["div_sweeps"] > 100:
        print(k, "t=%.4f dt=%.2e div=%d den=%d" % (g.time, dt, i["div_sweeps"], i["density_sweeps"]), flush=True)
print("done", k, len(mass))
