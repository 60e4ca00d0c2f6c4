"""Debug: resting dam break at 1M on the GPU up to just before its failure, then ONE step on the GPU and on the CPU oracle from
the same state."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import asph_b200 as A
from bench import uniform_params
spacing = 1.122e-3; stop = int(sys.argv[1]) if len(sys.argv) > 1 else 499
params = uniform_params(A)
scene = A.SceneConfig.dam_break(spacing, pos=(-0.95, -1 + 0.5 * spacing), size=(0.7, 1.8), fill=1.0)
pos, vel, mass = A.scene_particles(scene)
b = A.scene_boundary(scene, "AnalyticOverestimate")
g = A.FluidSimulation(params, pos, vel, mass, b)
for k in range(stop):
    g.single_step()
state = (g.get_field("position"), g.get_field("velocity"), g.get_field("mass"))
print("state at step", stop, "vmax", np.abs(state[1]).max(), "x range", state[0].min(0), state[0].max(0))
olib = A.load_library(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "liboracle_f32.so"))
for name, lib in (("gpu", None), ("oracle", olib)):
    s = A.FluidSimulation(params, state[0], state[1], state[2], b, lib=lib) if lib else A.FluidSimulation(params, *state, b)
    for k in range(3):
        try:
            dt = s.single_step_without_adaptivity()
            i = s.step_info()
            rho = s.get_field("density"); p = s.get_field("pressure"); aii = s.get_field("aii"); nc = s.get_field("neighbor_count")
            print(name, "step", k, "dt %.3e div %d den %d rho [%.4f, %.4f] p max %.4g aii min %.4g ncount max %d" %
                  (dt, i["div_sweeps"], i["density_sweeps"], rho.min(), rho.max(), np.abs(p).max(), aii.min(), nc.max()), flush=True)
        except Exception as e:
            print(name, "step", k, "FAILED", str(e)[:90], flush=True)
            try:
                rho = s.get_field("density"); aii = s.get_field("aii"); nc = s.get_field("neighbor_count"); p = s.get_field("pressure")
                print("   rho [%.4f, %.4f] aii [%.4g, %.4g] ncount max %d nonfinite p %d" % (rho.min(), rho.max(), aii.min(), aii.max(), nc.max(), (~np.isfinite(p)).sum()))
                bad = np.where(~np.isfinite(p))[0][:5]
                print("   bad particles", bad, state[0][bad], "aii", aii[bad], "rho", rho[bad], "nc", nc[bad])
            except Exception as e2:
                print("   (fields unavailable)", str(e2)[:80])
            break
    s.close()
