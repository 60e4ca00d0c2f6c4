"""Exploration: how do variants of the dam-break initial condition behave (sweeps per step, stability)?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import asph_b200 as A
from bench import uniform_params, SPACING_C2
y0 = float(sys.argv[1]); fill = float(sys.argv[2]); steps = int(sys.argv[3]); width = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
params = uniform_params(A)
scene = A.SceneConfig.dam_break(SPACING_C2, pos=(-0.95, y0), size=(width, 1.8), fill=fill)
pos, vel, mass = A.scene_particles(scene)
sim = A.FluidSimulation(params, pos, vel, mass, A.scene_boundary(scene, "AnalyticOverestimate"), counters_enabled=True)
rows = []
for k in range(steps):
    c0 = sim.counters()["simulation-step"][0]
    try:
        dt = sim.single_step()
    except Exception as e:
        print("FAILED at step", k, str(e)[:80]); break
    i = sim.step_info()
    rows.append((k, sim.time, dt, i["div_sweeps"], i["density_sweeps"], sim.counters()["simulation-step"][0] - c0))
print(f"y0={y0} fill={fill} n={len(mass)} steps={len(rows)}")
den = np.array([r[4] for r in rows]); div = np.array([r[3] for r in rows]); ms = np.array([r[5] for r in rows])
for a in range(0, len(rows), max(1, len(rows) // 12)):
    b = min(len(rows), a + max(1, len(rows) // 12))
    print(f"steps {a:4d}-{b:4d}: t={rows[b-1][1]:.4f} dt={rows[b-1][2]:.2e} div avg {div[a:b].mean():6.1f} max {div[a:b].max():4d}  den avg {den[a:b].mean():6.1f} max {den[a:b].max():4d}  ms avg {ms[a:b].mean():.2f}")
