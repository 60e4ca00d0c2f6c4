"""Multi-GPU parity worker (launched by torchrun from test_dist_gpu.py, one rank per GPU).

Every rank steps its slab of a small dam break through the distributed handle; rank 0 also steps the whole scene on a
plain single-GPU handle.  The N-GPU fields, gathered into reference order, must equal the 1-GPU fields: same dt and
sweep counts, positions / velocities / densities / pressures within fp32 summation-order noise.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import asph_b200 as A

    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    solver = sys.argv[2] if len(sys.argv) > 2 else "HybridDFSPH"
    spacing = float(sys.argv[3]) if len(sys.argv) > 3 else 0.006
    mode = sys.argv[4] if len(sys.argv) > 4 else "random"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    params = A.SimulationParams.from_yaml(os.path.join(ROOT, "configs", "default-config.yaml")).replace(
        merging=False, sharing=False, splitting=False, level_estimation_method="None", pressure_solver_method=solver)
    if mode in ("adaptive", "levelset"):
        return adaptive(A, dist, rank, world, steps, spacing, mode)
    # "random": seeded random velocities => compression somewhere from the first step, the pressure solver iterates
    #           (the recipe of test_single_step_uniform); few steps, because an SPH impact amplifies rounding noise
    # "drift":  the block moves to the right in free fall => particles migrate between slabs every step
    # "columns": the benchmark's weak-scaling scene in small (bench.py `dam_break`): separate fluid columns with empty
    #           space between them, slab faces through the middle of a column, random velocities as in "random"
    if mode == "columns":
        tank = 2.0 * world
        spans = [(-tank / 2 + 0.05, 0.3)] + [(-tank / 2 + 2.0 * k - 0.3, 0.6) for k in range(1, world)] + [(tank / 2 - 0.35, 0.3)]
        scene = A.SceneConfig({"boundary": {"type": "box", "width": tank, "height": 2.0},
                               "blocks": [{"pos": [x0, -0.98], "size": [w, 0.5], "spacing": spacing, "volume_fill_ratio": 0.93,
                                           "velocity": [0, 0]} for x0, w in spans]})
    else:
        scene = A.SceneConfig.dam_break(spacing, pos=(-0.9, -0.6), size=(1.2, 0.5))
    pos, vel, mass = A.scene_particles(scene)
    rng = np.random.default_rng(7)
    if mode in ("random", "columns"):
        vel = (rng.standard_normal(vel.shape) * 0.05).astype(np.float32)
    else:
        vel = (rng.standard_normal(vel.shape) * 0.002).astype(np.float32)
        vel[:, 0] += np.float32(0.8)
    n_global = len(mass)
    lo, hi = A.share_range(n_global, rank, world)
    boundary = A.scene_boundary(scene, "AnalyticOverestimate")
    d = A.DistributedFluidSimulation(params, pos[lo:hi], vel[lo:hi], mass[lo:hi], np.arange(lo, hi, dtype=np.uint32), n_global,
                                     boundary, counters_enabled=True)
    single = None
    if rank == 0:
        single = A.FluidSimulation(params, pos, vel, mass, boundary)
    owned0 = d.num_fluid_particles()
    report = {"world": world, "steps": steps, "n_global": d.n_global, "solver": solver, "rows": []}
    ok = True
    for k in range(steps):
        dt = d.single_step()
        info = d.step_info()
        if single is not None:
            dt1 = single.single_step()
            i1 = single.step_info()
            row = {"step": k, "dt": dt, "dt1": dt1, "div": info["div_sweeps"], "div1": i1["div_sweeps"],
                   "den": info["density_sweeps"], "den1": i1["density_sweeps"]}
            report["rows"].append(row)
            # dt is an exact min over the particles, but of velocities that differ in their last bits between the two
            # runs (different grid origins => different summation orders): equal to an ulp, not bit for bit
            if abs(dt - dt1) > 2e-6 * dt1:
                ok = False
    owned = d.num_fluid_particles()
    fields = {}
    for name in ("position", "velocity", "density", "pressure", "mass"):
        fields[name] = d.gather_field(name)
    if rank == 0:
        err = {}
        for name, scale in (("position", 2.0), ("velocity", None), ("density", 1.0), ("pressure", None), ("mass", None)):
            a, b = fields[name].astype(np.float64), single.get_field(name).astype(np.float64)
            s = scale if scale is not None else max(np.abs(b).max(), 1e-30)
            err[name] = float(np.abs(a - b).max() / s)
        report["err"] = err
        report["sweeps_equal"] = all(r["div"] == r["div1"] and r["den"] == r["den1"] for r in report["rows"])
        report["dt_equal"] = ok
        report["max_div_sweeps"] = max(r["div"] for r in report["rows"])
        report["max_density_sweeps"] = max(r["den"] for r in report["rows"])
    counts = [None] * world
    dist.all_gather_object(counts, (owned0, owned))
    if rank == 0:
        report["owned_first"] = [c[0] for c in counts]
        report["owned"] = [c[1] for c in counts]
        report["rows"] = report["rows"][-3:]
        print("DIST_REPORT " + json.dumps(report))
    d.close()
    if single is not None:
        single.close()
    dist.barrier()
    dist.destroy_process_group()


def adaptive(A, dist, rank, world, steps, spacing, mode):
    """BASELINE configs[2] recipe in small (tests/test_gpu_parity.py::test_adaptive_dam_break_mid_size): level set and, in mode
    "adaptive", share / merge / split across the slabs.  Every step: the particle count, the level sweeps and the resampling
    statistics of the N-GPU run equal the 1-GPU run's; at the end the fields, gathered by reference index, agree."""
    r_f = float(np.sqrt(0.93 / np.pi) * spacing)
    params = A.SimulationParams.from_yaml(os.path.join(ROOT, "configs", "default-config.yaml")).replace(
        particle_radius_fine=r_f, particle_radius_base=4.0 * r_f, maximum_surface_distance=0.2)
    if mode == "levelset":
        params = params.replace(merging=False, sharing=False, splitting=False)
    split = A.load_split_patterns_from_file()
    scene = A.SceneConfig.dam_break(spacing, pos=(-0.9, -0.9), size=(1.2, 0.8))
    pos, vel, mass = A.scene_particles(scene)
    n_global = len(mass)
    lo, hi = A.share_range(n_global, rank, world)
    boundary = A.scene_boundary(scene, "AnalyticOverestimate")
    d = A.DistributedFluidSimulation(params, pos[lo:hi], vel[lo:hi], mass[lo:hi], np.arange(lo, hi, dtype=np.uint32), n_global,
                                     boundary, counters_enabled=True, split_patterns=split)
    single = A.FluidSimulation(params, pos, vel, mass, boundary, split) if rank == 0 else None
    keys = ("div_sweeps", "density_sweeps", "level_sweeps", "n_shared", "n_merged", "n_split_parents")
    report = {"world": world, "steps": steps, "n_global": n_global, "mode": mode, "rows": [], "mismatch": []}
    for k in range(steps):
        dt = d.single_step()
        info = d.step_info()
        n_now = d.num_global_particles()
        if single is not None:
            dt1 = single.single_step()
            i1 = single.step_info()
            row = {"step": k, "n": n_now, "n1": single.num_fluid_particles(), "dt": dt, "dt1": dt1}
            for key in keys:
                row[key] = [int(info[key]), int(i1[key])]
                if int(info[key]) != int(i1[key]):
                    report["mismatch"].append((k, key, int(info[key]), int(i1[key])))
            if row["n"] != row["n1"]:
                report["mismatch"].append((k, "n", row["n"], row["n1"]))
            if abs(dt - dt1) > 2e-6 * dt1:
                report["mismatch"].append((k, "dt", dt, dt1))
            report["rows"].append(row)
    fields = {name: d.gather_field(name) for name in ("position", "velocity", "mass", "level")}
    owned = d.num_fluid_particles()
    counts = [None] * world
    dist.all_gather_object(counts, owned)
    if rank == 0:
        err = {}
        for name, scale in (("position", 2.0), ("velocity", None), ("mass", None), ("level", 1.0)):
            a, b = fields[name].astype(np.float64), single.get_field(name).astype(np.float64)
            if a.shape != b.shape:
                err[name] = float("inf")
                continue
            s = scale if scale is not None else max(np.abs(b).max(), 1e-30)
            err[name] = float(np.abs(a - b).max() / s)
        report["err"] = err
        report["owned"] = counts
        report["n_end"] = [int(len(fields["mass"])), single.num_fluid_particles()]
        report["merged_total"] = int(sum(r["n_merged"][1] for r in report["rows"]))
        report["split_total"] = int(sum(r["n_split_parents"][1] for r in report["rows"]))
        report["shared_total"] = int(sum(r["n_shared"][1] for r in report["rows"]))
        report["rows"] = report["rows"][-2:]
        print("DIST_REPORT " + json.dumps(report))
    d.close()
    if single is not None:
        single.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
