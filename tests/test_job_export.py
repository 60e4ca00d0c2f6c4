"""Batch export jobs (`image JOB.yaml`, platform/desktop/animation/mod.rs:28-288).  Host logic
only: the jobs are stepped by the oracle library handed in from here; the product CLI binds libasph_b200.so."""
import glob
import json
import os

import numpy as np
import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "configs", "default-config.yaml")
SCENE = os.path.join(ROOT, "configs", "default-scene.yaml")
SMALL_SCENE = {"boundary": {"type": "box", "width": 2, "height": 2},
               "blocks": [{"pos": [-0.9, -0.9], "size": [0.4, 0.5], "spacing": 0.03, "volume_fill_ratio": 0.93, "velocity": [0, 0]}]}


def _write_jobs(tmp_path, jobs):
    p = tmp_path / "jobs.yaml"
    p.write_text(yaml.safe_dump(jobs))
    return str(p)


def test_still_job_matches_a_manual_run(asph, oracle32, tmp_path):
    job = {"time": 0.0055, "config_path": CFG, "scene": SMALL_SCENE, "png_file": "still.png", "output_stats": True,
           "visualization_params": {"visualized_attribute": "Density"}, "title": "t",
           "update_attributes": {"max_dt": 0.002, "init_boundary_handler": "AnalyticUnderestimate"}}
    path = _write_jobs(tmp_path, [job])
    out = tmp_path / "out"
    (m,) = asph.export_simulation_image([path], oracle32, out_dir=str(out))
    assert m["finished"] and m["steps"] == 3 and not m["video"]  # 3 * 0.002 >= 0.0055
    assert sorted(os.listdir(out)) == ["still.png.job.json", "still.png.stat", "still.png.vtk"]
    assert json.load(open(out / "still.png.job.json"))["visualization_params"]["visualized_attribute"] == "Density"
    assert "simulation-step" in (out / "still.png.stat").read_text()
    # the same three steps by hand: physics, adaptivity, physics, adaptivity, physics (the export comes before the
    # third step's resampling, animation/mod.rs:138-273)
    params, scene = asph.resolve_job(asph.load_job_file(path)[0], str(tmp_path))
    sim = asph.init_fluid_sim(params, scene, asph.load_split_patterns_from_file(), lib=oracle32)
    for k in range(3):
        dt = sim.single_step_without_adaptivity(params)
        if k < 2:
            sim.single_step_adaptivity(params, dt)
    snap = asph.read_vtk_file(str(out / "still.png.vtk"))
    assert np.array_equal(snap["position"], sim.get_field("position"))
    assert np.array_equal(snap["density"], sim.get_field("density"))
    sim.close()


def test_video_job_frames_are_interpolated(asph, oracle32, tmp_path):
    job = {"time": 0.005, "video_start_time": 0, "video_fps": 1000, "config_path": CFG, "scene": SMALL_SCENE,
           "png_file": "clip.mp4", "visualization_params": {"visualized_attribute": "Velocity"},
           "update_attributes": {"max_dt": 0.002, "merging": False, "sharing": False, "splitting": False}}
    path = _write_jobs(tmp_path, [job])
    (m,) = asph.export_simulation_image([path], oracle32, out_dir=str(tmp_path))
    # steps end at t = 0.002, 0.004, 0.006 > 0.005; frames every 1 ms from 0: t = 0, ..., 0.006 (fp32 accumulation)
    assert m["finished"] and m["steps"] == 3 and m["frames"] in (6, 7)
    frames = sorted(glob.glob(str(tmp_path / "clip.mp4.frames" / "file-*.vtk")))
    assert len(frames) == m["frames"]
    params, scene = asph.resolve_job(asph.load_job_file(path)[0], str(tmp_path))
    x0, _, _ = asph.scene_particles(scene)
    sim = asph.init_fluid_sim(params, scene, None, lib=oracle32)
    sim.single_step_without_adaptivity(params)
    x1 = sim.get_field("position")
    sim.close()
    f0 = asph.read_vtk_file(frames[0])["position"]
    f1 = asph.read_vtk_file(frames[1])["position"]
    assert np.array_equal(f0, x0)  # interpolation weight 0: the positions before the first step
    assert np.allclose(f1, 0.5 * (x0 + x1), atol=1e-6)  # t = 1 ms, halfway through the first step
    series = json.load(open(tmp_path / "clip.mp4.frames" / "frames.vtk.series"))
    assert len(series["files"]) == m["frames"]


def test_job_errors(asph, oracle32, tmp_path):
    base = {"time": 0.001, "config_path": CFG, "png_file": "x.png", "visualization_params": {"visualized_attribute": "Density"}}
    with pytest.raises(asph.JobError, match="either 'scene' or 'scene_file'"):
        asph.export_simulation_image([_write_jobs(tmp_path, [dict(base)])], oracle32, out_dir=str(tmp_path))
    with pytest.raises(asph.JobError, match="Not both"):
        asph.export_simulation_image([_write_jobs(tmp_path, [dict(base, scene=SMALL_SCENE, scene_file=SCENE)])], oracle32, out_dir=str(tmp_path))
    with pytest.raises(asph.JobError, match="not able to find attribute nope"):
        asph.export_simulation_image([_write_jobs(tmp_path, [dict(base, scene=SMALL_SCENE, update_attributes={"nope": 1})])], oracle32, out_dir=str(tmp_path))
    with pytest.raises(asph.JobError, match="missing field `png_file`"):
        asph.load_job_file(_write_jobs(tmp_path, [{k: v for k, v in base.items() if k != "png_file"}]))
    with pytest.raises(asph.JobError, match="REACHED END BEFORE EXPORT"):  # never fires before the export in a still job...
        job = dict(base, scene=SMALL_SCENE, panic_on_end=True, time=0.0005, update_attributes={"max_dt": 0.002})
        asph.export_simulation_image([_write_jobs(tmp_path, [job])], oracle32, out_dir=str(tmp_path))  # ...unless a step overshoots `time`


def test_cli_image_subcommand(asph, oracle32, tmp_path, capsys):
    import importlib
    cli = importlib.import_module("adaptive-sph_b200.cli")
    job = {"time": 1.0, "config_path": CFG, "scene_file": SCENE, "png_file": "a.png",
           "visualization_params": {"visualized_attribute": "Density"}}
    path = _write_jobs(tmp_path, [job, dict(job, png_file="b.png")])
    rc = cli.main(["image", path, "--out-dir", str(tmp_path / "o"), "--only", "1", "--max-steps", "2", "-q"], lib=oracle32)
    assert rc == 0
    assert "1 job(s), 0 reached their export time" in capsys.readouterr().out
    assert os.listdir(tmp_path / "o") == ["b.png.job.json"]


REF_MEDIA = "/root/reference/media"


@pytest.mark.skipif(not os.path.isdir(REF_MEDIA), reason="the reference tree is only mounted in the build container")
def test_every_media_job_of_the_reference_resolves(asph):
    """All job files the reference ships parse, and their overrides resolve against its parameter files unmodified."""
    n_jobs = 0
    modes = set()
    stale = {}
    for path in sorted(glob.glob(os.path.join(REF_MEDIA, "*.yaml"))):
        with open(path) as f:
            doc = yaml.safe_load(f)
        if not (isinstance(doc, list) and doc and isinstance(doc[0], dict) and "png_file" in doc[0]):
            continue  # a scene file
        try:
            for job in asph.load_job_file(path):
                assert job.unknown_keys == []
                params, scene = asph.resolve_job(job, REF_MEDIA)
                assert asph.scene_particle_count(scene) > 0
                modes.add((params["pressure_solver_method"], params["support_length_estimation"], params["operator_discretization"]))
                n_jobs += 1
        except asph.JobError as e:
            stale[os.path.basename(path)] = str(e)
    assert n_jobs >= 74 and len(modes) >= 4
    # Job files the reference's own current schema rejects the same way: four predate `visualization_params` (a
    # mandatory field of ImageExportConfig, animation/mod.rs:49), one overrides an Option field that is absent from
    # default-config.yaml (`.unwrap_or_else(|| panic!("not able to find attribute {}"))`, animation/mod.rs:96)
    assert sorted(stale) == ["constant-field.yaml", "density.yaml", "distance-to-neighbor.yaml", "render-test.yaml", "surface-distance.yaml"]
    assert all("visualization_params" in v or "fill_stash_with" in v for v in stale.values())


@pytest.mark.gpu
def test_jobs_on_the_gpu_match_the_oracle(asph, cuda_lib, oracle32, tmp_path):
    """A still and a video job through the product library: same steps, same frame count, snapshot positions within
    1e-6 of the domain size of the oracle's."""
    still = {"time": 0.0055, "config_path": CFG, "scene": SMALL_SCENE, "png_file": "still.png", "output_stats": True,
             "visualization_params": {"visualized_attribute": "Density"}, "update_attributes": {"max_dt": 0.002}}
    clip = {"time": 0.005, "video_start_time": 0, "video_fps": 1000, "config_path": CFG, "scene": SMALL_SCENE, "png_file": "clip.mp4",
            "visualization_params": {"visualized_attribute": "Velocity"}, "update_attributes": {"max_dt": 0.002}}
    path = _write_jobs(tmp_path, [still, clip])
    mg = asph.export_simulation_image([path], cuda_lib, out_dir=str(tmp_path / "gpu"))
    mo = asph.export_simulation_image([path], oracle32, out_dir=str(tmp_path / "cpu"))
    for a, b in zip(mg, mo):
        assert a["backend"] == "cuda-sm100a" and b["backend"] == "oracle-f32"
        for k in ("finished", "steps", "frames", "particles_first_step", "particles_end"):
            assert a[k] == b[k], (k, a, b)
    files = ["still.png.vtk"] + [os.path.join("clip.mp4.frames", os.path.basename(f))
                                 for f in sorted(glob.glob(str(tmp_path / "cpu" / "clip.mp4.frames" / "file-*.vtk")))]
    for f in files:
        g = asph.read_vtk_file(str(tmp_path / "gpu" / f)); o = asph.read_vtk_file(str(tmp_path / "cpu" / f))
        assert np.abs(g["position"] - o["position"]).max() <= 2e-6, f
        assert np.allclose(g["mass"], o["mass"], rtol=1e-6), f
    assert "simulation-step" in (tmp_path / "gpu" / "still.png.stat").read_text()
