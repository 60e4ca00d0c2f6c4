"""Headless `run <config> <scene>` — the reference's desktop command line without the window.

Mirrors the clap definition of platform/desktop/main_loop.rs:25-103 for the `run` sub-command:
  run SIMULATION_CONFIG SCENE_CONFIG [-s/--max-seconds S] [-c/--overwrite-config-file F] [-p/--statistics-enabled]
                                     [-w/--statistics-path F]
and its flow (main_loop.rs:105-189, 209-358): read YAML -> optional key-wise overwrite -> init_simulation_params ->
load split patterns -> init_fluid_sim -> loop { single_step } until the simulated time reaches --max-seconds.
`image JOB.yaml...` runs the reference's batch export jobs (animation/mod.rs:28-288) with VTK snapshots in place of the
Cairo/ffmpeg rendering (export_jobs.py); `generate-split-patterns` (offline optimiser) is out of scope.
Added: --max-steps, --dump state.npz.  The CUDA library is the only backend; `main(argv, lib=...)` lets a test harness
bind another library exporting the same C ABI.
"""
import argparse
import os
import sys
import time

import numpy as np

from .binding import StatisticsRecorder, init_fluid_sim, load_library
from .params import SimulationParams
from .scene import SceneConfig, init_simulation_params
from .split_patterns import load_split_patterns_from_file
from .vtk import VtkExporter, init_fluid_sim_from_vtk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_parser():
    ap = argparse.ArgumentParser(prog="asph_b200", description="Adaptive SPH step loop on B200 (headless)")
    sub = ap.add_subparsers(dest="command", required=True)
    run = sub.add_parser("run", help="Run a simulation")
    run.add_argument("simulation_config", metavar="SIMULATION_CONFIG")
    run.add_argument("scene_config", metavar="SCENE_CONFIG")
    run.add_argument("-s", "--max-seconds", type=float, default=None, help="Stops the simulation after n simulated seconds")
    run.add_argument("-c", "--overwrite-config-file", default=None)
    run.add_argument("-p", "--statistics-enabled", action="store_true")
    run.add_argument("-w", "--statistics-path", default=None)
    run.add_argument("--max-steps", type=int, default=None)
    run.add_argument("--split-patterns", default=None, help="default: ./split-patterns.yaml if present, else the shipped file")
    run.add_argument("--dump", default=None, help="write the final state (position, velocity, mass) to this .npz")
    run.add_argument("--vtk-dir", default=None, help="write VTK snapshots (vtk_exporter.rs format) and a .vtk.series index into this folder")
    run.add_argument("--vtk-every", type=int, default=1, help="snapshot every n-th step")
    run.add_argument("--restart-vtk", default=None, help="start from this VTK snapshot instead of the scene's blocks (the scene still gives the boundary)")
    run.add_argument("-q", "--quiet", action="store_true")
    img = sub.add_parser("image", help="Run batch export jobs (VTK snapshots instead of rendered images)")
    img.add_argument("job_files", metavar="JOB_FILE", nargs="+")
    img.add_argument("--out-dir", default=None, help="default: next to the job file, as the reference does")
    img.add_argument("--only", type=int, action="append", default=None, help="run only job number K of each file (repeatable)")
    img.add_argument("--max-steps", type=int, default=None, help="stop every job after n steps (regression runs)")
    img.add_argument("--split-patterns", default=None)
    img.add_argument("-q", "--quiet", action="store_true")
    sub.add_parser("generate-split-patterns", help="not available in the headless B200 build")
    return ap


def main(argv=None, lib=None):
    args = build_parser().parse_args(argv)
    if args.command == "image":
        from .export_jobs import JobError, export_simulation_image
        if lib is None:
            lib = load_library()
        try:
            done = export_simulation_image(args.job_files, lib, out_dir=args.out_dir, only=args.only, max_steps=args.max_steps,
                                           quiet=args.quiet, split_patterns_path=args.split_patterns)
        except JobError as e:
            print(f"job failed: {e}", file=sys.stderr)
            return 1
        print(f"{len(done)} job(s), {sum(1 for m in done if m['finished'])} reached their export time")
        return 0
    if args.command != "run":
        print(f"`{args.command}` is out of scope of this build (offline pattern optimiser)", file=sys.stderr)
        return 2
    if args.max_seconds is None and args.max_steps is None:
        print("headless run needs --max-seconds or --max-steps", file=sys.stderr)
        return 2
    params = SimulationParams.from_yaml(args.simulation_config, overwrite_path=args.overwrite_config_file)
    scene = SceneConfig.from_yaml(args.scene_config)
    params = init_simulation_params(params, scene)
    sp_path = args.split_patterns or ("./split-patterns.yaml" if os.path.exists("./split-patterns.yaml") else None)  # main_loop.rs:225
    split = load_split_patterns_from_file(sp_path)
    if lib is None:
        lib = load_library()  # raises when libasph_b200.so is missing: there is no CPU fallback
    stats_on = args.statistics_enabled or args.statistics_path is not None
    if args.restart_vtk:
        sim = init_fluid_sim_from_vtk(params, scene, args.restart_vtk, split, counters_enabled=stats_on, lib=lib)
    else:
        sim = init_fluid_sim(params, scene, split, counters_enabled=stats_on, lib=lib)
    vtk = VtkExporter(args.vtk_dir, "my-sph") if args.vtk_dir else None  # main_loop.rs:256
    rec = StatisticsRecorder()
    step = 0
    t0 = time.perf_counter()
    while True:
        if args.max_seconds is not None and sim.time >= args.max_seconds:
            break
        if args.max_steps is not None and step >= args.max_steps:
            break
        ts = time.perf_counter()
        if vtk is not None and step % max(1, args.vtk_every) == 0:
            # the split form of the step, as the reference's exporter uses it (animation/mod.rs:138-273): the per-step
            # fields of a snapshot describe the particle set of the physics step, before resampling changes it
            dt = sim.single_step_without_adaptivity(params)
            vtk.add_snapshot(sim.time, sim, params)
            sim.single_step_adaptivity(params, dt)
        else:
            dt = sim.single_step(params)
        info = sim.step_info()
        rec.record_step(info)
        step += 1
        if not args.quiet:
            print(f"step {step}: t={sim.time:.5f} dt={dt:.3e} n={info['n_particles_end']} div-iters={info['div_iterations']} "
                  f"density-iters={info['density_iterations']} shared={info['n_shared']} merged={info['n_merged']} "
                  f"split={info['n_split_parents']}  {1e3 * (time.perf_counter() - ts):.2f}ms")
    wall = time.perf_counter() - t0
    print(f"{step} steps, simulated {sim.time:.4f} s, {sim.num_fluid_particles()} particles, wall {wall:.2f} s, backend {sim.backend()}")
    if stats_on:
        text = rec.write_statistics(sim)
        if args.statistics_path:
            with open(args.statistics_path, "w") as f:
                f.write(text)
        else:
            print(text)
    if args.dump:
        np.savez_compressed(args.dump, position=sim.get_field("position"), velocity=sim.get_field("velocity"),
                            mass=sim.get_field("mass"))
    sim.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
