"""Multi-GPU arm of bench.py (N > 1, one rank per GPU under torchrun): slab decomposition along x, NCCL halos.

Weak scaling: the tank and the fluid block of BASELINE configs[1] are widened N-fold, so every GPU owns about 999 292
particles; the physics per column is unchanged (same spacing, same column height => same sweep counts as at N = 1).
Each rank generates only its own share of the lattice.  Timing: barrier + synchronize on both sides of the K timed
steps; per rank the CUDA-event time of the steps on the library's stream; the job's time is the MAX over ranks.
"""
import json
import os
import time

import numpy as np


def run(args, A, rank, world):
    import torch
    import torch.distributed as dist
    from bench import (METRIC, UNIT, SPACING_C2, ClockSampler, dam_break, peaks, pinned, preroll, uniform_params)

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = A.load_library()
    params = uniform_params(A)
    scene = dam_break(A, SPACING_C2, n_gpus=world)
    sim = A.DistributedFluidSimulation.from_scene(params, scene, counters_enabled=True, lib=lib, rank=rank, world=world, device=local)
    n_global = sim.n_global
    K, W = args.steps, args.warmup

    def fence():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    pre_steps = preroll(sim, args.preroll_time)   # every rank sees the same global dt, hence the same step count
    for _ in range(W):
        sim.single_step()
    sim.set_kernel_timing(4)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    fence()
    cnt0 = sim.counters()
    c0 = cnt0["simulation-step"][0]
    l0 = sim.kernel_launches()
    t0 = time.perf_counter()
    owned_steps, sweeps_div, sweeps_den = 0, 0, 0
    for _ in range(K):
        sim.single_step()
        info = sim.step_info()
        owned_steps += info["n_particles_begin"]
        sweeps_div += info["div_sweeps"]; sweeps_den += info["density_sweeps"]
    fence()
    wall = time.perf_counter() - t0
    cnt1 = sim.counters()
    dev_ms = cnt1["simulation-step"][0] - c0
    phases = {k: (cnt1[k][0] - cnt0[k][0]) / max(K, 1) for k in cnt1}
    launches = sim.kernel_launches() - l0
    clk = clocks.stop() if rank == 0 else None
    kt = sim.kernel_timing()
    sim.set_kernel_timing(0)

    # ---- e2e: host buffers in (this rank's owned particles), host buffers out, every step -----------------------
    n_own = sim.num_fluid_particles()
    cap = int(n_own * 1.25) + 65536
    hp, _a = pinned((cap, 2)); hv, _b = pinned((cap, 2)); hm, _c = pinned((cap,))
    op, _d = pinned((cap, 2)); ov, _e = pinned((cap, 2)); om, _f = pinned((cap,))
    sim.get_field("position", out=hp[:n_own]); sim.get_field("velocity", out=hv[:n_own]); sim.get_field("mass", out=hm[:n_own])
    fence()
    t1 = time.perf_counter()
    e2e_owned, h2d, d2h = 0, 0, 0
    for _ in range(K):
        sim.set_state(hp[:n_own], hv[:n_own], hm[:n_own])      # H2D of this step's inputs
        h2d += n_own * 20
        e2e_owned += n_own
        sim.single_step()
        n_own = sim.num_fluid_particles()                        # migration may have changed the owned set
        sim.get_field("position", out=op[:n_own]); sim.get_field("velocity", out=ov[:n_own]); sim.get_field("mass", out=om[:n_own])
        d2h += n_own * 20
        hp[:n_own] = op[:n_own]; hv[:n_own] = ov[:n_own]; hm[:n_own] = om[:n_own]
    fence()
    e2e_s = time.perf_counter() - t1

    # ---- reduce over ranks: sums of work, MAX of time --------------------------------------------------------------
    t = torch.tensor([float(owned_steps), float(launches), float(e2e_owned), float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    m = torch.tensor([dev_ms, wall * 1e3, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    owned = torch.tensor([float(n_own)], dtype=torch.float64, device="cuda")
    owned_all = [torch.zeros_like(owned) for _ in range(world)]
    dist.all_gather(owned_all, owned)
    if rank == 0:
        total_steps, total_launches, e2e_total, h2d_t, d2h_t = [float(x) for x in t.tolist()]
        dev_ms_max, wall_ms_max, e2e_ms_max = [float(x) for x in m.tolist()]
        value = total_steps / (dev_ms_max * 1e-3)
        peak, peak_src = peaks()
        n_rank = total_steps / max(K, 1) / world
        roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                "kernel": "k_jacobi on rank 0 (K15, 40 B/particle algorithmic over owned + ghost particles)", "peak_source": peak_src}
        if kt["jacobi_sweep"][1] > 0:
            ms_j = kt["jacobi_sweep"][0] / kt["jacobi_sweep"][1]
            roof["achieved"] = 40.0 * n_rank / (ms_j * 1e-3) / 1e9
            roof["frac"] = roof["achieved"] / peak
            roof["avg_launch_ms"] = ms_j
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms_max / max(K, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[1] widened {world}x: 2D dam-break, uniform h, {n_global} particles ({n_global // world} per GPU), "
                                   "HybridDFSPH, x-slab decomposition, NCCL ghost halos per pass; "
                                   f"state at t = {args.preroll_time} s (just after the block hits the floor: both pressure solves iterate)",
                       "preroll_steps": pre_steps, "preroll_time_s": args.preroll_time,
                       "particles": n_global, "owned_per_rank": [int(x.item()) for x in owned_all],
                       "l2": "working set per GPU (~400 MB) exceeds the 126 MB L2; no flush",
                       "avg_div_sweeps": sweeps_div / max(K, 1), "avg_density_sweeps": sweeps_den / max(K, 1),
                       "timing": "max over ranks of the CUDA-event time of the K steps on the library stream; barrier + synchronize on both sides",
                       "wall_ms_per_step": wall_ms_max / max(K, 1),
                       "phase_ms_per_step_rank0": phases,
                       "sweep_kernels_us_rank0": {k: (kt[k][0] / kt[k][1] * 1e3 if kt[k][1] else None) for k in ("accel_sweep", "jacobi_sweep", "neighbors", "sort_grid")}},
            "clocks": clk,
            "e2e": {"value": e2e_total / (e2e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_t / max(K, 1)),
                    "d2h_bytes_per_step": int(d2h_t / max(K, 1))},
            "gpu_launches": int(total_launches), "roofline": roof,
            "cpu_baseline": None,
        }
        print(json.dumps(out), flush=True)
    sim.close()
    dist.barrier()
    dist.destroy_process_group()
