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
