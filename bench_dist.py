"""Multi-GPU arm of bench.py (N > 1, one rank per GPU under torchrun): the SAME scene as at N = 1 — the north star's
adaptive 16 M-particle dam break — cut into N x-slabs of equal particle count (strong scaling).

Every rank generates its contiguous share of the initial lattice; the library migrates particles to their owner slabs,
exchanges the ghost zones once per step with NCCL, and inside the step everything that crosses a slab face travels over
NVLink peer memory from within the kernels: the Jacobi sweeps store border values straight into the neighbour GPU's
arrays, the persistent level-set and partner-search kernels mail theirs and meet in cross-GPU barriers.
The input preparation (untimed) is the one of bench.py: PREROLL_STEPS steps with the base radius ramped up.
Timing: barrier + synchronize on both sides of the K timed steps; per rank the CUDA-event time of its steps on the
library's stream; the job's time is the MAX over ranks; value = sum over the steps of the particles of the whole fluid
at the start of the step / that time.
"""
import json
import os
import time

import numpy as np


def run(args, A, rank, world):
    import torch
    import torch.distributed as dist
    from bench import (KERNEL_BYTES, METRIC, PREROLL_STEPS, SPACING, UNIT, ClockSampler, StepLog, adaptive_params, dam_break,
                       kernel_table, peaks, pinned, preroll_adaptive, workload_name)

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t_start = time.perf_counter()
    lib = A.load_library()
    base = adaptive_params(A, SPACING)
    scene = dam_break(A, SPACING)
    base = A.init_simulation_params(base, scene)
    split = A.load_split_patterns_from_file()
    K, W = args.steps, args.warmup

    def fence():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    sim = A.DistributedFluidSimulation.from_scene(base, scene, counters_enabled=True, lib=lib, rank=rank, world=world, device=local,
                                                  split_patterns=split)
    n0 = sim.n_global
    t_pre = time.perf_counter()
    preroll_adaptive(sim, A, base, SPACING)
    for _ in range(W):
        sim.single_step(base)
    t_pre = time.perf_counter() - t_pre

    # ---- device-resident arm ----------------------------------------------------------------------------------------
    sim.set_kernel_timing(1)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    fence()
    cnt0 = sim.counters()
    l0 = sim.kernel_launches()
    log = StepLog()
    dev_ms, wall = 0.0, time.perf_counter()
    for _ in range(K):
        cb = sim.counters()["simulation-step"][0]
        sim.single_step(base)
        dev_ms += sim.counters()["simulation-step"][0] - cb
        log.add(sim)   # n = this rank's owned particles at the start of the step
    fence()
    wall = time.perf_counter() - wall
    clk = clocks.stop() if rank == 0 else None
    cnt1 = sim.counters()
    phases = {k: (cnt1[k][0] - cnt0[k][0]) / max(K, 1) for k in cnt1}
    launches = sim.kernel_launches() - l0
    kt = sim.kernel_timing()
    sim.set_kernel_timing(0)

    # ---- e2e: host buffers in (this rank's owned particles), host buffers out, every step --------------------------------
    n_own = sim.num_fluid_particles()
    cap = int(n_own * 1.5) + 65536
    hp, _a = pinned((cap, 2)); hv, _b = pinned((cap, 2)); hm, _c = pinned((cap,))
    sim.get_field("position", out=hp[:n_own]); sim.get_field("velocity", out=hv[:n_own]); sim.get_field("mass", out=hm[:n_own])
    fence()
    e2e_log = StepLog()
    h2d = d2h = 0
    t1 = time.perf_counter()
    for _ in range(K):
        sim.set_state(hp[:n_own], hv[:n_own], hm[:n_own])      # H2D of this step's inputs
        h2d += n_own * 20
        sim.single_step(base)
        e2e_log.add(sim)
        n_own = sim.num_fluid_particles()                        # migration and resampling change the owned set
        if n_own > cap:
            raise RuntimeError("owned particle count outgrew the host buffers")
        sim.get_field("position", out=hp[:n_own]); sim.get_field("velocity", out=hv[:n_own]); sim.get_field("mass", out=hm[:n_own])
        d2h += n_own * 20
    fence()
    e2e_s = time.perf_counter() - t1

    # ---- reduce over ranks: sums of work, MAX of time ----------------------------------------------------------------
    t = torch.tensor([float(sum(log.n)), float(launches), float(sum(e2e_log.n)), float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    m = torch.tensor([dev_ms, wall * 1e3, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    owned_all = [None] * world
    dist.all_gather_object(owned_all, int(n_own))
    if rank == 0:
        total_steps, total_launches, e2e_total, h2d_t, d2h_t = [float(x) for x in t.tolist()]
        dev_ms_max, wall_ms_max, e2e_ms_max = [float(x) for x in m.tolist()]
        value = total_steps / (dev_ms_max * 1e-3)
        peak, peak_src = peaks()
        sm = log.summary()
        n_rank = float(sum(log.n)) / max(K, 1)
        tab = kernel_table(kt, K, n_rank, peak)
        single = {k: v for k, v in tab.items() if k in KERNEL_BYTES}
        dom = max(single, key=lambda k: single[k]["ms_per_step"]) if single else None
        roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src}
        if dom:
            roof.update({"kernel": KERNEL_BYTES[dom][1] + " on rank 0 (algorithmic bytes over its owned particles; the launch includes its waits for the neighbour GPUs)",
                         "achieved": single[dom]["achieved_gbs"], "frac": single[dom]["frac"], "avg_launch_ms": single[dom]["avg_launch_ms"],
                         "share_of_step": single[dom]["ms_per_step"] / (dev_ms / max(K, 1)), "particles_per_launch": n_rank})
        s_div, s_den = sm.get("avg_div_sweeps", 0.0), sm.get("avg_density_sweeps", 0.0)
        b_step = 388.0 + 68.0 * (s_div + s_den)
        sm_global = dict(sm)
        sm_global["particles_first"] = None  # per-rank figures; the global counts are below
        sm_global["particles_last"] = None
        e2e_work = e2e_log.summary()
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms_max / max(K, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict({"workload": workload_name(n0) + f"; {world} x-slabs of equal particle count, one per GPU",
                            "particles_initial": n0, "particles_per_step": total_steps / max(K, 1), "owned_per_rank": owned_all,
                            "preroll_steps": PREROLL_STEPS, "preroll_wall_s": t_pre,
                            "l2": "working set per GPU exceeds the 126 MB L2 up to 8 GPUs; no flush",
                            "timing": "max over ranks of the CUDA-event time of the K steps on the library stream; barrier + synchronize on both sides",
                            "wall_ms_per_step": wall_ms_max / max(K, 1), "phase_ms_per_step_rank0": phases, "kernels_rank0": tab,
                            "switches": {k: os.environ[k] for k in ("ASPH_BULK", "ASPH_DIST_P2P", "ASPH_P2P_EDGE_FIRST", "ASPH_SWEEP_GRID", "ASPH_BENCH_SPACING", "ASPH_BENCH_PREROLL") if k in os.environ}},
                           **sm_global),
            "clocks": clk,
            "e2e": {"value": e2e_total / (e2e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_t / max(K, 1)),
                    "d2h_bytes_per_step": int(d2h_t / max(K, 1)), "ms_per_step": e2e_ms_max / max(K, 1),
                    "work": {k: e2e_work[k] for k in e2e_work if k.startswith("avg_")},
                    "same_work_as_device_arm": all(abs(e2e_work.get(k, 0) - sm.get(k, 0)) <= 1.0 for k in ("avg_div_sweeps", "avg_density_sweeps", "avg_level_sweeps"))},
            "gpu_launches": int(total_launches), "roofline": roof,
            "step_roofline": {"bytes_per_particle_step": b_step, "achieved": b_step * value / 1e9, "frac": b_step * value / 1e9 / (peak * world), "unit": "GB/s",
                              "peak": peak * world},
            "cpu_baseline": None, "bench_wall_s": time.perf_counter() - t_start,
        }
        print(json.dumps(out), flush=True)
    sim.close()
    dist.barrier()
    dist.destroy_process_group()
