"""Multi-GPU probe (torchrun): the benchmark scene widened world-fold, stepped until failure; per-step sweeps and ownership."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
import asph_b200 as A
from bench import uniform_params, dam_break, SPACING_C2
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
spacing = float(sys.argv[2]) if len(sys.argv) > 2 else SPACING_C2
params = uniform_params(A)
scene = dam_break(A, spacing, n_gpus=world)
sim = A.DistributedFluidSimulation.from_scene(params, scene, counters_enabled=True, rank=rank, world=world, device=local)
for k in range(steps):
    try:
        dt = sim.single_step()
        ok = 1
    except Exception as e:
        ok = 0
        err = str(e)[:100]
    i = sim.step_info() if ok else {}
    t = torch.tensor([float(sim.num_fluid_particles()), float(ok)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    if rank == 0 and (k % 10 == 0 or not ok or i.get("density_sweeps", 0) > 200 or i.get("div_sweeps", 0) > 200):
        print(f"step {k} t={sim.time:.4f} owned_total={int(t[0].item())} of {sim.n_global} ok={int(t[1].item())}/{world} " +
              (f"dt={dt:.2e} div={i['div_sweeps']} den={i['density_sweeps']}" if ok else err), flush=True)
    if t[1].item() < world:
        break
sim.close()
dist.barrier()
dist.destroy_process_group()
