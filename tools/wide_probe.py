"""Single-GPU run of the benchmark scene widened n-fold (what bench.py gives N GPUs): sweeps per step until failure."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import asph_b200 as A
from bench import uniform_params, dam_break, SPACING_C2
n_gpus = int(sys.argv[1]); steps = int(sys.argv[2])
params = uniform_params(A)
scene = dam_break(A, SPACING_C2, n_gpus=n_gpus, kind="wide")
pos, vel, mass = A.scene_particles(scene)
sim = A.FluidSimulation(params, pos, vel, mass, A.scene_boundary(scene, "AnalyticOverestimate"))
for k in range(steps):
    try:
        dt = sim.single_step()
    except Exception as e:
        print("step", k, "FAILED", str(e)[:90], flush=True); break
    i = sim.step_info()
    if k % 10 == 0 or i["density_sweeps"] > 200 or i["div_sweeps"] > 200:
        print(f"step {k} t={sim.time:.4f} n={len(mass)} dt={dt:.2e} div={i['div_sweeps']} den={i['density_sweeps']}", flush=True)
