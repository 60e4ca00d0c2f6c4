"""Kernel-time probe: one step of the profiling workload (random velocities => the solves iterate) with every sweep timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import asph_b200 as A
from bench import uniform_params, dam_break, SPACING_C2
spacing = float(sys.argv[1]) if len(sys.argv) > 1 else SPACING_C2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
params = uniform_params(A).replace(max_iters=60)
scene = dam_break(A, spacing)
pos, vel, mass = A.scene_particles(scene)
rng = np.random.default_rng(5)
vel = (rng.standard_normal(vel.shape) * 0.02).astype(np.float32)
sim = A.FluidSimulation(params, pos, vel, mass, A.scene_boundary(scene, "AnalyticOverestimate"), counters_enabled=True)
sim.single_step()
sim.set_kernel_timing(1)
for k in range(steps):
    try:
        sim.single_step()
    except Exception as e:
        print("step failed:", str(e)[:60]); break
kt = sim.kernel_timing()
i = sim.step_info()
out = {k: (v[0] / v[1] * 1e3 if v[1] else None) for k, v in kt.items()}
print("steps", steps, "n", len(mass), "sweeps", i["div_sweeps"], i["density_sweeps"],
      " ".join(f"{k}={v:.1f}us" for k, v in out.items() if v is not None))
