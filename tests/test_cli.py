"""`run <config> <scene>` command line (platform/desktop/main_loop.rs:25-103), driven with the CPU oracle backend."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "asph_b200.py"), *args], capture_output=True, text=True, timeout=600)


def test_run_oracle_backend_with_statistics(tmp_path):
    stats = tmp_path / "stats.txt"
    over = tmp_path / "over.yaml"
    over.write_text("max_dt: 0.003\n")
    out = _run("run", os.path.join(ROOT, "configs", "default-config.yaml"), os.path.join(ROOT, "configs", "default-scene.yaml"),
               "-s", "0.0089", "-c", str(over), "-w", str(stats), "--backend", "oracle")
    assert out.returncode == 0, out.stderr
    assert "3 steps" in out.stdout  # 3 * 0.003 >= 0.0089
    text = stats.read_text()
    for label in ("simulation-step", "neighborhood", "level-estimation", "div-solver", "density-solver", "adaptivity", "particle-count", "dt:"):
        assert label in text


def test_unknown_overwrite_key_is_an_error(tmp_path):
    over = tmp_path / "over.yaml"
    over.write_text("not_a_field: 1\n")
    out = _run("run", os.path.join(ROOT, "configs", "default-config.yaml"), os.path.join(ROOT, "configs", "default-scene.yaml"),
               "--max-steps", "1", "-c", str(over), "--backend", "oracle")
    assert out.returncode != 0 and "not able to find attribute" in out.stderr


def test_out_of_scope_subcommands():
    assert _run("image").returncode == 2
