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
    if mode.startswith("resample"):
        return resample(A, dist, rank, world, steps, mode)
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


def resample(A, dist, rank, world, seed, mode):
    """single_step_adaptivity across the slabs on prescribed inputs (the recipe of test_resampling_phase_parity): a jittered
    lattice with masses spread over all five size classes of a prescribed level field, so that donors and receivers sit on
    both sides of every slab face.  After one physics step (lists, ghosts) both runs get the SAME level field and run the
    resampling phase alone: same statistics, bit-identical masses, by reference index."""
    import ctypes as C
    phase = mode.split(":", 1)[1] if ":" in mode else "share+merge"
    rng = np.random.default_rng(seed)
    sp = np.float32(0.01)
    blk = dict(pos=(np.float32(-0.8), np.float32(-0.5)), size=(np.float32(1.601), np.float32(0.801)), spacing=sp,
               volume_fill_ratio=np.float32(0.93), velocity=(np.float32(0), np.float32(0)))
    pos, vel, mass = A.add_fluid_block(blk)
    pos = (pos + rng.uniform(-0.25, 0.25, pos.shape).astype(np.float32) * sp).astype(np.float32)
    vel = (rng.standard_normal(pos.shape) * 0.01).astype(np.float32)
    mass = (mass * rng.uniform(0.3, 2.6, mass.shape)).astype(np.float32)
    level = np.minimum((-(0.35 - pos[:, 1]) * 0.5).astype(np.float32), np.float32(0.0))
    r0 = float(np.sqrt(0.93e-4 / np.pi))
    params = A.SimulationParams.from_yaml(os.path.join(ROOT, "configs", "default-config.yaml")).replace(
        particle_radius_fine=r0, particle_radius_base=2.0 * r0, maximum_surface_distance=0.2,
        sharing="share" in phase, merging="merge" in phase, splitting="split" in phase,
        max_dt=1e-5, max_iters=3)  # the physics step only has to build the lists and the ghosts: the masses are far from rest
    split = A.load_split_patterns_from_file()
    scene = A.SceneConfig({"boundary": {"type": "box", "width": 2.0, "height": 2.0}, "blocks": []})
    boundary = A.scene_boundary(scene, "AnalyticOverestimate")
    n_global = len(mass)
    order = np.argsort(pos[:, 0], kind="stable")     # the shares are x-slabs; the reference index stays the lattice index
    lo, hi = A.share_range(n_global, rank, world)
    mine = np.sort(order[lo:hi])
    d = A.DistributedFluidSimulation(params, pos[mine], vel[mine], mass[mine], mine.astype(np.uint32), n_global, boundary,
                                     split_patterns=split)
    single = A.FluidSimulation(params, pos, vel, mass, boundary, split) if rank == 0 else None
    dt = d.single_step_without_adaptivity()
    lp = level.ctypes.data_as(C.POINTER(C.c_float))
    step_number = 2 if "merge" in phase else 3
    assert d.lib.asph_set_level(d._h, lp, n_global) == 0
    d.lib.asph_set_step_number(d._h, step_number)
    d.single_step_adaptivity(dt=0.002)
    info = d.step_info()
    report = {"world": world, "mode": mode, "seed": seed, "n_global": n_global, "mismatch": [], "rounds": [d.adapt_rounds(), None]}
    n_now = d.num_global_particles()
    fields = {name: d.gather_field(name) for name in ("mass", "position", "velocity")}
    if single is not None:
        dt1 = single.single_step_without_adaptivity()
        assert single.lib.asph_set_level(single._h, lp, n_global) == 0
        single.lib.asph_set_step_number(single._h, step_number)
        single.single_step_adaptivity(dt=0.002)
        i1 = single.step_info()
        report["rounds"][1] = single.adapt_rounds()
        for key in ("n_shared", "n_merged", "n_split_parents"):
            report[key] = [int(info[key]), int(i1[key])]
            if int(info[key]) != int(i1[key]):
                report["mismatch"].append((key, int(info[key]), int(i1[key])))
        report["n"] = [n_now, single.num_fluid_particles()]
        if n_now != single.num_fluid_particles():
            report["mismatch"].append(("n", n_now, single.num_fluid_particles()))
        else:
            m1 = single.get_field("mass")
            bad = np.nonzero(fields["mass"] != m1)[0]
            report["mass_bits_differ"] = int(len(bad))
            report["first_bad"] = [int(b) for b in bad[:8]]
            report["pos_maxdiff"] = float(np.abs(fields["position"] - single.get_field("position")).max())
            report["vel_maxdiff"] = float(np.abs(fields["velocity"] - single.get_field("velocity")).max())
        report["dt"] = [dt, dt1]
        print("DIST_REPORT " + json.dumps(report))
        single.close()
    d.close()
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
        if os.environ.get("ASPH_DIST_TRACE"):  # per-step distance of the two runs (diagnostics; a gather per step)
            gm, gl, gx = d.gather_field("mass"), d.gather_field("level"), d.gather_field("position")
            if single is not None:
                sm_, sl_, sx_ = single.get_field("mass"), single.get_field("level"), single.get_field("position")
                if gm.shape == sm_.shape:
                    bad = np.nonzero(gm != sm_)[0]
                    report.setdefault("trace", []).append({"step": k, "mass_diff": int(len(bad)), "first_bad": [int(b) for b in bad[:6]],
                                                           "level_maxdiff": float(np.abs(gl - sl_).max()), "pos_maxdiff": float(np.abs(gx - sx_).max())})
                else:
                    report.setdefault("trace", []).append({"step": k, "shape": [int(gm.shape[0]), int(sm_.shape[0])]})
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
        report["mass_total"] = [float(fields["mass"].astype(np.float64).sum()), float(single.get_field("mass").astype(np.float64).sum())]
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
