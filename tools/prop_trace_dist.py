"""Development probe, several GPUs (torchrun): per-sweep cycle counts of k_propagate<PEER> on rank 0
(library built with `make EXTRA_level=-DASPH_PROP_TRACE`).

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/prop_trace_dist.py
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asph_b200 as A
import bench


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = A.load_library()
    base = bench.adaptive_params(A, bench.SPACING)
    scene = bench.dam_break(A, bench.SPACING)
    base = A.init_simulation_params(base, scene)
    sim = A.DistributedFluidSimulation.from_scene(base, scene, counters_enabled=True, lib=lib, rank=rank, world=world, device=local,
                                                  split_patterns=A.load_split_patterns_from_file())
    bench.preroll_adaptive(sim, A, base, bench.SPACING)
    for _ in range(3):
        sim.single_step(base)
    buf = np.zeros((12, 512), dtype=np.uint64)
    ptr = buf.ctypes.data_as(C.POINTER(C.c_ulonglong))
    lib.asph_debug_prop_trace(ptr, 1)
    sim.single_step(base)
    lib.asph_debug_prop_trace(ptr, 0)
    sw = int(sim.step_info()["level_sweeps"])
    if rank == 0:
        print("sweeps", sw, "warps", int(buf[5, 1]))
        print(" t   front border push_max push_mean  sync1   mailed  barrier  imported  sync2   (cycles since the sweep began, block 0)")
        for t in range(1, sw + 1):
            if t <= 8 or t % 10 == 0:
                print(f"{t:3d} {int(buf[0, t]):7d} {int(buf[10, t]):5d} {int(buf[1, t]):8d} {int(buf[2, t]) / max(1, int(buf[5, t])):9.0f} "
                      f"{int(buf[4, t]):7d} {int(buf[6, t]):7d} {int(buf[7, t]):8d} {int(buf[8, t]):8d} {int(buf[9, t]):7d}")
        r = slice(1, sw + 1)
        names = {0: "front", 10: "border", 1: "push_max", 4: "sync1", 6: "mailed", 7: "barrier", 8: "imported", 9: "sync2"}
        print("means:", {v: round(float(buf[k, r].mean()), 1) for k, v in names.items()})
    sim.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
