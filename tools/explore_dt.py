"""Exploration: resting dam break at 1M with max_dt scaled to the resolution: stability (density range) and sweeps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import asph_b200 as A
from bench import uniform_params
spacing = 1.122e-3; max_dt = float(sys.argv[1]); steps = int(sys.argv[2]); y_gap = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
params = uniform_params(A).replace(max_dt=max_dt)
scene = A.SceneConfig.dam_break(spacing, pos=(-0.95, -1 + y_gap * spacing), size=(0.7, 1.8), fill=1.0)
pos, vel, mass = A.scene_particles(scene)
g = A.FluidSimulation(params, pos, vel, mass, A.scene_boundary(scene, "AnalyticOverestimate"), counters_enabled=True)
acc = []
for k in range(steps):
    c0 = g.counters()["simulation-step"][0]
    try:
        dt = g.single_step_without_adaptivity()
    except Exception as e:
        print("FAILED at step", k, str(e)[:80], flush=True); break
    i = g.step_info()
    acc.append((i["div_sweeps"], i["density_sweeps"], g.counters()["simulation-step"][0] - c0))
    if (k + 1) % (steps // 10) == 0:
        a = np.array(acc); acc = []
        rho = g.get_field("density"); x = g.get_field("position")
        print("step %5d t=%.4f dt=%.2e div avg %.1f max %d den avg %.1f max %d ms %.2f rho [%.3f, %.3f] xmin %.5f ymin %.5f" %
              (k + 1, g.time, dt, a[:, 0].mean(), a[:, 0].max(), a[:, 1].mean(), a[:, 1].max(), a[:, 2].mean(), rho.min(), rho.max(), x[:, 0].min(), x[:, 1].min()), flush=True)
