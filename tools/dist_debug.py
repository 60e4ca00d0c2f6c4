"""Debug helper: per-step, per-field comparison of the N-GPU run against the 1-GPU run (torchrun)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import asph_b200 as A
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
params = A.SimulationParams.from_yaml(os.path.join(ROOT, "configs", "default-config.yaml")).replace(
    merging=False, sharing=False, splitting=False, level_estimation_method="None")
scene = A.SceneConfig.dam_break(0.006, pos=(-0.99, -0.99), size=(1.2, 0.5), fill=1.0)
d = A.DistributedFluidSimulation.from_scene(params, scene)
single = None
if rank == 0:
    pos, vel, mass = A.scene_particles(scene)
    single = A.FluidSimulation(params, pos, vel, mass, A.scene_boundary(scene, "AnalyticOverestimate"))
for k in range(steps):
    dt = d.single_step(); info = d.step_info()
    f = {n: d.gather_field(n) for n in ("position", "velocity", "density", "aii", "ppe_source_term", "pressure", "pressure_accel", "neighbor_count", "h")}
    if rank == 0:
        dt1 = single.single_step(); i1 = single.step_info()
        print(f"step {k}: dt {dt} {dt1} div {info['div_sweeps']} {i1['div_sweeps']} den {info['density_sweeps']} {i1['density_sweeps']}")
        bad_any = any(np.abs(a.astype(np.float64) - single.get_field(n).astype(np.float64)).max() > 1e-4 * max(np.abs(single.get_field(n)).max(), 1e-30) for n, a in f.items())
        if not bad_any:
            continue
        for n, a in f.items():
            b = single.get_field(n)
            diff = np.abs(a.astype(np.float64) - b.astype(np.float64))
            if diff.ndim > 1: diff = diff.max(axis=1)
            w = int(diff.argmax())
            print(f"   {n:16s} max|d|={diff.max():.3e} scale={np.abs(b).max():.3e} at gid {w} x={f['position'][w]} nbad={(diff > 1e-4 * max(np.abs(b).max(), 1e-30)).sum()}")
d.close()
dist.barrier(); dist.destroy_process_group()
