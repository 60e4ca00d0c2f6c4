"""The native host program (adaptive-sph_b200/host/asph_run, C++): YAML subset parser, SimulationParams / SceneConfig /
SplitPatterns loaders, the `run` command line of platform/desktop/main_loop.rs:25-189.  Checked against PyYAML and the Python
host logic, and end to end with the oracle library handed in through --lib (the program's default is the CUDA library)."""
import glob
import os
import subprocess
import json

import numpy as np
import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "adaptive-sph_b200", "host")
EXE = os.path.join(HOST, "asph_run")
CFG = os.path.join(ROOT, "configs", "default-config.yaml")
SCENE = os.path.join(ROOT, "configs", "default-scene.yaml")
ORACLE = os.path.join(ROOT, "oracle", "liboracle_f32.so")


@pytest.fixture(scope="session")
def exe(oracle32):  # oracle32: makes sure oracle/liboracle_f32.so is built
    subprocess.check_call(["make", "-C", HOST], stdout=subprocess.DEVNULL)
    return EXE


def _run(exe, *args, **kw):
    return subprocess.run([exe, *args], capture_output=True, text=True, timeout=600, **kw)


def _same(py, cpp, where="$"):
    """PyYAML's typed tree against the C++ parse (scalars as text)."""
    if isinstance(py, dict):
        assert isinstance(cpp, dict) and [str(k) for k in py] == list(cpp), where
        for k in py:
            _same(py[k], cpp[str(k)], f"{where}.{k}")
    elif isinstance(py, list):
        assert isinstance(cpp, list) and len(py) == len(cpp), where
        for i, (a, b) in enumerate(zip(py, cpp)):
            _same(a, b, f"{where}[{i}]")
    elif py is None:
        assert cpp is None, where
    elif isinstance(py, bool):
        assert cpp in ("true", "false", "True", "False") and (cpp.lower() == "true") == py, where
    elif isinstance(py, (int, float)):
        assert float(cpp) == float(py), (where, py, cpp)
    else:
        assert cpp == py, (where, py, cpp)


def _yaml_files():
    files = sorted(glob.glob(os.path.join(ROOT, "configs", "*.yaml"))) + [os.path.join(ROOT, "adaptive-sph_b200", "data", "split-patterns.yaml")]
    if os.path.isdir("/root/reference"):  # only in the build container
        files += sorted(glob.glob("/root/reference/*.yaml")) + sorted(glob.glob("/root/reference/media/*.yaml"))
    return files


def test_yaml_subset_parser_agrees_with_pyyaml(exe):
    files = _yaml_files()
    assert len(files) >= 7
    for path in files:
        out = _run(exe, "yaml-dump", path)
        assert out.returncode == 0, (path, out.stderr)
        with open(path) as f:
            _same(yaml.load(f, Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader)), json.loads(out.stdout), os.path.basename(path))


def test_params_struct_matches_the_python_loader(asph, exe, tmp_path):
    import ctypes as C
    over = tmp_path / "over.yaml"
    over.write_text("max_dt: 0.003\npressure_solver_method: IISPH\nlevel_estimation_method: None\nmerging: false\n")
    for args, kw in (((CFG,), {}), ((CFG, str(over)), {"overwrite_path": str(over)}),
                     ((os.path.join(ROOT, "configs", "default-config-web.yaml"),), {})):
        out = _run(exe, "params-dump", *args)
        assert out.returncode == 0, out.stderr
        p = asph.SimulationParams.from_yaml(args[0], **kw)
        assert bytes.fromhex(out.stdout.strip()) == bytes(C.string_at(C.addressof(p.c), C.sizeof(p.c)))
    bad = tmp_path / "bad.yaml"
    bad.write_text("not_a_field: 1\n")
    out = _run(exe, "params-dump", CFG, str(bad))
    assert out.returncode == 1 and "not able to find attribute not_a_field" in out.stderr
    missing = tmp_path / "missing.yaml"
    missing.write_text("".join(l for l in open(CFG) if not l.startswith("gravity")))
    out = _run(exe, "params-dump", str(missing))
    assert out.returncode == 1 and "missing field(s) gravity" in out.stderr
    wrong = tmp_path / "wrong.yaml"
    wrong.write_text("viscosity_type: Honey\n")
    out = _run(exe, "params-dump", CFG, str(wrong))
    assert out.returncode == 1 and "unknown variant `Honey`" in out.stderr


def _read_dump(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"ASPHDUMP"
    n = int(np.frombuffer(raw[8:16], np.uint64)[0])
    a = np.frombuffer(raw[16:], np.float32)
    assert len(a) == 5 * n
    return a[:2 * n].reshape(n, 2), a[2 * n:4 * n].reshape(n, 2), a[4 * n:]


def test_scene_particles_match_add_fluid_block(asph, exe, tmp_path):
    for name in ("default-scene.yaml", "default-scene-web.yaml", "ratio-stress-test-scene.yaml", "motivation-scene2.yaml"):
        path = os.path.join(ROOT, "configs", name)
        dump = tmp_path / (name + ".bin")
        out = _run(exe, "scene-dump", path, str(dump))
        assert out.returncode == 0, out.stderr
        pos, vel, mass = _read_dump(dump)
        p, v, m = asph.scene_particles(asph.SceneConfig.from_yaml(path))
        assert np.array_equal(pos, p) and np.array_equal(vel, v) and np.array_equal(mass, m), name


def test_run_matches_the_python_host_bit_for_bit(asph, oracle32, exe, tmp_path):
    over = tmp_path / "over.yaml"
    over.write_text("init_boundary_handler: AnalyticUnderestimate\nmax_dt: 0.004\n")
    dump, stats = tmp_path / "state.bin", tmp_path / "stats.txt"
    out = _run(exe, "run", CFG, SCENE, "--max-steps", "6", "-c", str(over), "--dump", str(dump), "-w", str(stats), "--lib", ORACLE, "-q")
    assert out.returncode == 0, out.stderr
    assert "6 steps" in out.stdout and "backend oracle-f32" in out.stdout
    params = asph.SimulationParams.from_yaml(CFG, overwrite_path=str(over))
    scene = asph.SceneConfig.from_yaml(SCENE)
    params = asph.init_simulation_params(params, scene)
    sim = asph.init_fluid_sim(params, scene, asph.load_split_patterns_from_file(), lib=oracle32)
    for _ in range(6):
        sim.single_step(params)
    pos, vel, mass = _read_dump(dump)
    assert np.array_equal(pos, sim.get_field("position")) and np.array_equal(vel, sim.get_field("velocity"))
    assert np.array_equal(mass, sim.get_field("mass"))
    sim.close()
    text = stats.read_text()
    for label in ("simulation-step", "neighborhood", "level-estimation", "div-solver", "density-solver", "adaptivity", "particle-count", "dt:"):
        assert label in text


def test_run_stops_at_max_seconds_and_reports_errors(exe, tmp_path):
    over = tmp_path / "over.yaml"
    over.write_text("max_dt: 0.003\n")
    out = _run(exe, "run", CFG, SCENE, "-s", "0.0089", "-c", str(over), "--lib", ORACLE, "-q")
    assert out.returncode == 0 and "3 steps" in out.stdout  # 3 * 0.003 >= 0.0089
    assert _run(exe, "run", CFG, SCENE).returncode == 2                      # needs --max-seconds or --max-steps
    out = _run(exe, "run", CFG, SCENE, "--max-steps", "1", "--lib", str(tmp_path / "nope.so"))
    assert out.returncode == 1 and "no CPU fallback" in out.stderr
    assert _run(exe, "image").returncode == 2


@pytest.mark.gpu
def test_run_on_the_gpu(exe, tmp_path):
    dump = tmp_path / "state.bin"
    out = _run(exe, "run", CFG, SCENE, "--max-steps", "5", "--dump", str(dump), "-p")
    assert out.returncode == 0, out.stderr
    assert "5 steps" in out.stdout and "backend cuda-sm100a" in out.stdout
    pos, vel, mass = _read_dump(dump)
    assert len(mass) > 1035 and np.all(np.isfinite(pos))


def _same_snapshot(asph, a, b, exact=True):
    sa, sb = asph.read_vtk_file(a), asph.read_vtk_file(b)
    assert set(sa) == set(sb), (sorted(sa), sorted(sb))
    for k in sa:
        if k == "distances" or not exact:
            assert np.allclose(sa[k], sb[k], rtol=1e-6, atol=1e-7), k
        else:
            assert np.array_equal(sa[k], sb[k]), k


def test_vtk_snapshots_equal_the_python_exporter(asph, oracle32, exe, tmp_path):
    """`run --vtk-dir`: same files, same arrays as the Python VtkExporter on the same run (polygon boundary: lines,
    distances and lambda arrays included)."""
    import importlib
    cli = importlib.import_module("adaptive-sph_b200.cli")
    over = tmp_path / "over.yaml"
    over.write_text("init_boundary_handler: AnalyticUnderestimate\n")
    out = _run(exe, "run", CFG, SCENE, "--max-steps", "3", "-c", str(over), "--vtk-dir", str(tmp_path / "cpp"), "--lib", ORACLE, "-q")
    assert out.returncode == 0, out.stderr
    assert cli.main(["run", CFG, SCENE, "--max-steps", "3", "-c", str(over), "--vtk-dir", str(tmp_path / "py"), "-q"], lib=oracle32) == 0
    names = sorted(os.listdir(tmp_path / "py"))
    assert sorted(os.listdir(tmp_path / "cpp")) == names == ["my-sph-00001.vtk", "my-sph-00002.vtk", "my-sph-00003.vtk", "my-sph.vtk.series"]
    for n in names[:3]:
        _same_snapshot(asph, str(tmp_path / "cpp" / n), str(tmp_path / "py" / n))
        assert os.path.getsize(tmp_path / "cpp" / n) == os.path.getsize(tmp_path / "py" / n)
    sc, sp = json.load(open(tmp_path / "cpp" / names[3])), json.load(open(tmp_path / "py" / names[3]))
    assert [f["name"] for f in sc["files"]] == [f["name"] for f in sp["files"]]
    assert np.allclose([f["time"] for f in sc["files"]], [f["time"] for f in sp["files"]], rtol=1e-6)


def test_image_jobs_equal_the_python_exporter(asph, oracle32, exe, tmp_path):
    small = {"boundary": {"type": "box", "width": 2, "height": 2},
             "blocks": [{"pos": [-0.9, -0.9], "size": [0.4, 0.5], "spacing": 0.03, "volume_fill_ratio": 0.93, "velocity": [0, 0]}]}
    still = {"time": 0.0055, "config_path": CFG, "scene": small, "png_file": "still.png", "output_stats": True,
             "visualization_params": {"visualized_attribute": "Density"}, "update_attributes": {"max_dt": 0.002}}
    clip = {"time": 0.005, "video_start_time": 0, "video_fps": 1000, "config_path": CFG, "scene_file": SCENE, "png_file": "sub/clip.mp4",
            "visualization_params": {"visualized_attribute": "Velocity"}, "update_attributes": {"max_dt": 0.002, "splitting": False}}
    jobs = tmp_path / "jobs.yaml"
    jobs.write_text(yaml.safe_dump([still, clip]))
    out = _run(exe, "image", str(jobs), "--out-dir", str(tmp_path / "cpp"), "--lib", ORACLE, "-q")
    assert out.returncode == 0, out.stderr
    assert "2 job(s), 2 reached their export time" in out.stdout
    man = asph.export_simulation_image([str(jobs)], oracle32, out_dir=str(tmp_path / "py"), log=lambda *a: None)
    _same_snapshot(asph, str(tmp_path / "cpp" / "still.png.vtk"), str(tmp_path / "py" / "still.png.vtk"))
    fc = sorted(glob.glob(str(tmp_path / "cpp" / "sub" / "clip.mp4.frames" / "file-*.vtk")))
    fp = sorted(glob.glob(str(tmp_path / "py" / "sub" / "clip.mp4.frames" / "file-*.vtk")))
    assert len(fc) == len(fp) == man[1]["frames"]
    for a, b in zip(fc, fp):
        _same_snapshot(asph, a, b)
    mc = json.load(open(tmp_path / "cpp" / "still.png.job.json"))
    assert mc["finished"] and mc["steps"] == man[0]["steps"] and mc["particles_end"] == man[0]["particles_end"]
    assert "simulation-step" in (tmp_path / "cpp" / "still.png.stat").read_text()
    # the reference's panics
    bad = tmp_path / "bad.yaml"
    bad.write_text(yaml.safe_dump([dict(still, update_attributes={"nope": 1})]))
    out = _run(exe, "image", str(bad), "--out-dir", str(tmp_path / "x"), "--lib", ORACLE)
    assert out.returncode == 1 and "not able to find attribute nope" in out.stderr
    bad.write_text(yaml.safe_dump([dict(still, scene_file=SCENE)]))
    out = _run(exe, "image", str(bad), "--out-dir", str(tmp_path / "x"), "--lib", ORACLE)
    assert out.returncode == 1 and "Not both" in out.stderr


def test_yaml_subset_parser_edge_cases(exe, tmp_path):
    """Constructs the shipped files do not exercise together: quoted scalars with '#' and ':', comments after values,
    sequences at the indentation of their key, nested flow collections, empty values, '- - x' items, blank lines."""
    doc = '''# leading comment
title: "Particle count #p: now"   # trailing comment
path: 'a: b'
empty:
nothing: ~
flag: true
num: -3.5e-05
list_same_indent:
- 1
- [2, 3, [4, 5]]
- {a: 1, b: [x, y], c: {d: e}}
nested:
  deeper:
    - - 0.5
      - -0.25
    - - 1
      - 2

  after_blank: ok
jobs:
  - time: 3
    update_attributes:
      viscosity: 0.001
      level_estimation_method: None
    scene:
      blocks:
        - pos: [0, 1]
          size: [1, 1]
  - time: 4
'''
    p = tmp_path / "edge.yaml"
    p.write_text(doc)
    out = _run(exe, "yaml-dump", str(p))
    assert out.returncode == 0, out.stderr
    _same(yaml.safe_load(doc), json.loads(out.stdout))
    bad = tmp_path / "bad.yaml"
    bad.write_text("a: [1, 2\n")
    out = _run(exe, "yaml-dump", str(bad))
    assert out.returncode == 1 and "unterminated flow collection" in out.stderr


def test_restart_from_a_snapshot_continues_bit_for_bit(asph, exe, tmp_path):
    """--restart-vtk: 3 steps, snapshot after the third, then 2 more from the snapshot == 5 steps in one run.  The snapshot is
    taken between the physics part and the resampling of its step, so the continued run starts with that step's resampling
    missing; the comparison therefore runs with resampling off (the Python harness has the same test with it on, through
    set-ups the native command line does not expose)."""
    over = tmp_path / "over.yaml"
    over.write_text("merging: false\nsharing: false\nsplitting: false\n")
    whole, part, cont = tmp_path / "whole.bin", tmp_path / "part", tmp_path / "cont.bin"
    assert _run(exe, "run", CFG, SCENE, "--max-steps", "5", "-c", str(over), "--dump", str(whole), "--lib", ORACLE, "-q").returncode == 0
    assert _run(exe, "run", CFG, SCENE, "--max-steps", "3", "-c", str(over), "--vtk-dir", str(part), "--lib", ORACLE, "-q").returncode == 0
    out = _run(exe, "run", CFG, SCENE, "--max-steps", "2", "-c", str(over), "--restart-vtk", str(part / "my-sph-00003.vtk"), "--dump", str(cont),
               "--lib", ORACLE, "-q")
    assert out.returncode == 0, out.stderr
    a, b = _read_dump(whole), _read_dump(cont)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    snap = asph.read_vtk_file(str(part / "my-sph-00003.vtk"))
    assert len(snap["mass"]) == len(a[2])


def test_yaml_subset_parser_on_generated_documents(exe, tmp_path):
    """Random nested documents in both of PyYAML's dump styles (block and flow) parse like PyYAML parses them."""
    from hypothesis import HealthCheck, given, settings, strategies as st
    keys = st.text("abcdefghijklmnopqrstuvwxyz_", min_size=1, max_size=8)
    words = st.text("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ", min_size=1, max_size=10).filter(
        lambda s: s.lower() not in ("null", "true", "false", "yes", "no", "on", "off", "y", "n", "nan", "inf"))
    tricky = st.sampled_from(["a b", "x: y", "# c", "a #b", "k:v", "- d", "[e", "1.5.2", "Particle count #p", "../default-config.yaml"])
    scalars = st.one_of(tricky, st.integers(-10 ** 6, 10 ** 6), st.floats(-1e6, 1e6, allow_nan=False, allow_infinity=False, width=32), st.booleans(), words, st.none())
    docs = st.recursive(scalars, lambda ch: st.one_of(st.lists(ch, min_size=1, max_size=4), st.dictionaries(keys, ch, min_size=1, max_size=4)), max_leaves=20)
    path = tmp_path / "gen.yaml"

    @settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(doc=st.dictionaries(keys, docs, min_size=1, max_size=5), flow=st.sampled_from([False, None]))
    def check(doc, flow):
        text = yaml.safe_dump(doc, default_flow_style=flow, width=10 ** 6)
        path.write_text(text)
        out = _run(exe, "yaml-dump", str(path))
        assert out.returncode == 0, (text, out.stderr)
        _same(yaml.safe_load(text), json.loads(out.stdout))

    check()
