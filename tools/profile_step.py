"""Profiling workload: the BASELINE configs[1] particle set with seeded random velocities, so that both Jacobi solves
iterate from the first step (kernel cost does not depend on the values).  Run under ncu (see profiles/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import asph_b200 as A
from bench import uniform_params, dam_break, SPACING_C2
spacing = float(sys.argv[1]) if len(sys.argv) > 1 else SPACING_C2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
adaptive = len(sys.argv) > 3 and sys.argv[3] == "adaptive"
params = uniform_params(A)
scene = dam_break(A, spacing)
pos, vel, mass = A.scene_particles(scene)
rng = np.random.default_rng(5)
vel = (rng.standard_normal(vel.shape) * 0.02).astype(np.float32)
if adaptive:
    mass = (mass * rng.uniform(0.8, 1.6, mass.shape)).astype(np.float32)
sim = A.FluidSimulation(params, pos, vel, mass, A.scene_boundary(scene, "AnalyticOverestimate"), counters_enabled=True)
for k in range(steps):
    c0 = sim.counters()["simulation-step"][0]
    sim.single_step()
    i = sim.step_info()
    print(f"step {k}: n={len(mass)} div_sweeps={i['div_sweeps']} density_sweeps={i['density_sweeps']} ms={sim.counters()['simulation-step'][0] - c0:.3f}")
sim.close()
