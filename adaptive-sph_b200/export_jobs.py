"""Batch export jobs: the reference's `image <job.yaml> ...` command without the renderer.

A job file is a YAML list of `ImageExportConfig` records (platform/desktop/animation/mod.rs:28-57): a parameter file,
key-wise `update_attributes` overrides, an inline `scene` or a `scene_file`, the simulated `time` at which the still is
taken — or `video_start_time` / `video_fps` / `video_speed` for a sequence of frames up to `time` — and the output name.
`export_simulation_image` follows animation/mod.rs:59-288 step for step: the split form of the step
(`single_step_without_adaptivity` → export while `time_for_next_export <= time` → `single_step_adaptivity`), frame
positions interpolated linearly between the positions before and after the physics step, `panic_on_end`, and the
`.stat` file of `write_statistics` when `output_stats` is set.

What differs, because Cairo / ffmpeg rendering is out of scope: where the reference rasterises a PNG the
job writes a VTK snapshot in the layout of vtk_exporter.rs (`<png_file>.vtk`; frames of a video job go to
`<png_file>.frames/file-000000.vtk ...` plus a `.vtk.series` index), and a `<png_file>.job.json` manifest records what
was run (resolved overrides, visualised attribute, steps, particle counts).  Every `media/*.yaml` of the reference runs
unmodified as a regression / benchmark suite this way.  Outputs go to `out_dir` (default: the job file's directory, as in
the reference).
"""
import json
import os

import numpy as np
import yaml

from .binding import StatisticsRecorder, init_fluid_sim
from .params import SimulationParams, merge_overwrite
from .scene import SceneConfig, init_simulation_params
from .split_patterns import load_split_patterns_from_file
from .vtk import write_vtk_file

# fields of ImageExportConfig (animation/mod.rs:28-57); serde rejects nothing here (no deny_unknown_fields), but a
# mistyped key would silently do nothing, so unknown keys are reported
_KNOWN = {"time", "video_start_time", "video_fps", "video_speed", "zoom_out", "interpolated", "no_legend",
          "legend_text_right", "legend_only_min_max", "title", "config_path", "scene", "scene_file", "update_attributes",
          "visualization_params", "png_file", "output_stats", "panic_on_end", "export_when_mii_negative",
          "video_img_dir", "image_width", "image_height"}
_REQUIRED = ("time", "config_path", "visualization_params", "png_file")


class JobError(Exception):
    """A panic of the reference's exporter (bad job file, `panic_on_end`, negative interpolation)."""


class ImageExportConfig:
    """One record of a job file.  Attribute names are the reference's field names."""

    def __init__(self, mapping):
        for k in _REQUIRED:
            if k not in mapping:
                raise JobError(f"failed parsing export config file: missing field `{k}`")
        self.unknown_keys = sorted(set(mapping) - _KNOWN)
        f32 = lambda v: None if v is None else float(np.float32(v))
        self.time = f32(mapping["time"])
        self.video_start_time = f32(mapping.get("video_start_time"))
        self.video_fps = f32(mapping.get("video_fps"))
        self.video_speed = f32(mapping.get("video_speed"))
        self.title = mapping.get("title")
        self.config_path = str(mapping["config_path"])
        self.scene = mapping.get("scene")
        self.scene_file = mapping.get("scene_file")
        self.update_attributes = dict(mapping.get("update_attributes") or {})
        self.visualization_params = dict(mapping["visualization_params"] or {})
        if "visualized_attribute" not in self.visualization_params:
            raise JobError("failed parsing export config file: missing field `visualized_attribute`")
        self.png_file = str(mapping["png_file"])
        self.output_stats = mapping.get("output_stats")
        self.panic_on_end = mapping.get("panic_on_end")
        self.video_img_dir = mapping.get("video_img_dir")

    @property
    def is_video(self):
        return self.video_start_time is not None


def load_job_file(path):
    with open(path) as f:
        records = yaml.safe_load(f)
    if not isinstance(records, list):
        raise JobError("failed parsing export config file: expected a list of jobs")
    return [ImageExportConfig(r) for r in records]


def resolve_job(job, job_dir):
    """Parameter file + overrides + scene of one job (animation/mod.rs:75-104) → (SimulationParams, SceneConfig)."""
    with open(os.path.join(job_dir, job.config_path)) as f:
        mapping = yaml.safe_load(f)
    if (job.scene is None) == (job.scene_file is None):
        raise JobError("expected either 'scene' or 'scene_file'" + (". Not both!" if job.scene is not None else ""))
    if job.scene is not None:
        scene = SceneConfig(job.scene)
    else:
        scene = SceneConfig.from_yaml(os.path.join(job_dir, job.scene_file))
    try:
        mapping = merge_overwrite(mapping, job.update_attributes)  # unknown key: "not able to find attribute"
    except KeyError as e:
        raise JobError(str(e.args[0])) from None
    params = init_simulation_params(SimulationParams(mapping), scene)
    return params, scene


def run_job(job, job_dir, out_dir, lib, split_patterns=None, max_steps=None, quiet=True, log=print):
    """One job, animation/mod.rs:75-288.  Returns the manifest dictionary that is also written next to the output."""
    params, scene = resolve_job(job, job_dir)
    sim = init_fluid_sim(params, scene, split_patterns, counters_enabled=True, lib=lib)
    rec = StatisticsRecorder()
    f32 = np.float32
    video = job.is_video
    fps = f32(job.video_fps if job.video_fps is not None else 60.0)
    speed = f32(job.video_speed if job.video_speed is not None else 1.0)
    end = f32(job.time)
    next_export = f32(job.video_start_time) if video else end
    out_path = os.path.join(out_dir, job.png_file)
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    frames_dir = out_path + ".frames"
    frames = []
    if video:
        os.makedirs(frames_dir, exist_ok=True)
        for old in os.listdir(frames_dir):  # "re-create video image dir"
            if old.endswith(".vtk") or old.endswith(".vtk.series"):
                os.remove(os.path.join(frames_dir, old))
    steps = 0
    finished = False
    counts = []
    try:
        while not finished:
            if max_steps is not None and steps >= max_steps:
                break
            time_before = f32(sim.time)
            pos_before = sim.get_field("position") if video else None
            dt = sim.single_step_without_adaptivity(params)
            steps += 1
            info = sim.step_info()
            rec.record_step(info)
            counts.append(int(info["n_particles_begin"]))
            now = f32(sim.time)
            if job.panic_on_end is True and now > end:
                raise JobError(">>>>>>>>>>>> REACHED END BEFORE EXPORT <<<<<<<<<<<<")
            while next_export <= now:
                if video:
                    a = f32((next_export - time_before) / (now - time_before))
                    if a < 0:
                        raise JobError(f"negative interpolation {a} (export {next_export} between {now} and {time_before})")
                    assert a <= 1.0
                    pos = (a * sim.get_field("position") + (f32(1.0) - a) * pos_before).astype(np.float32)
                    name = "file-{:06d}.vtk".format(len(frames))
                    write_vtk_file(os.path.join(frames_dir, name), sim, params, positions=pos)
                    frames.append((name, str(next_export)))
                    next_export = f32(next_export + f32(1.0) / fps * speed)
                    if now > end:
                        finished = True
                        break
                else:
                    write_vtk_file(out_path + ".vtk", sim, params)
                    finished = True
                    break
            if finished:
                break
            sim.single_step_adaptivity(params, dt)
            if not quiet:
                log(f"  step {steps}: t={float(now):.5f} dt={dt:.3e} n={info['n_particles_end']}")
        if video:
            with open(os.path.join(frames_dir, "frames.vtk.series"), "w") as f:
                f.write('{\n"file-series-version": "1.0",\n"files": [')
                f.write(",".join('\n{{ "name": "{}", "time": {} }}'.format(n, t) for n, t in frames))
                f.write("\n]\n}")
        if job.output_stats is True:
            with open(out_path + ".stat", "w") as f:
                f.write(rec.write_statistics(sim))
        manifest = {
            "png_file": job.png_file, "title": job.title, "finished": bool(finished), "steps": steps,
            "simulated_time": float(sim.time), "export_time": float(end), "video": bool(video), "frames": len(frames),
            "particles_first_step": counts[0] if counts else int(sim.num_fluid_particles()),
            "particles_end": int(sim.num_fluid_particles()),
            "visualization_params": job.visualization_params, "update_attributes": job.update_attributes,
            "config_path": job.config_path, "scene_file": job.scene_file, "backend": sim.backend(),
            "unknown_job_keys": job.unknown_keys,
        }
        with open(out_path + ".job.json", "w") as f:
            json.dump(manifest, f, indent=1, default=str)
        return manifest
    finally:
        sim.close()


def export_simulation_image(job_paths, lib, out_dir=None, only=None, max_steps=None, quiet=True, split_patterns_path=None,
                            log=print):
    """`image` sub-command (animation/mod.rs:59): every job of every job file, in order.  `only` = indices of the jobs
    of each file to run (None: all).  Returns the list of manifests."""
    sp_path = split_patterns_path or ("./split-patterns.yaml" if os.path.exists("./split-patterns.yaml") else None)  # mod.rs:105
    split = load_split_patterns_from_file(sp_path)
    manifests = []
    for p in job_paths:
        path = os.path.realpath(p)
        job_dir = os.path.dirname(path)
        jobs = load_job_file(path)
        for k, job in enumerate(jobs):
            if only is not None and k not in only:
                continue
            log(f"{os.path.basename(path)}[{k}] -> {job.png_file}")
            manifests.append(run_job(job, job_dir, out_dir or job_dir, lib, split, max_steps=max_steps, quiet=quiet, log=log))
    return manifests
