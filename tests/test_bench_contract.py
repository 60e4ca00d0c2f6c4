"""bench.py's reference arm end to end on CPU (the oracle port on a bounded sample) and the JSON contract of its line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    # what torchrun does to its workers (the arm must undo it); a scene the CPU prepares in seconds (no GPU here)
    env = dict(os.environ, OMP_NUM_THREADS="1", ASPH_BENCH_SPACING="0.01", ASPH_BENCH_PREROLL="6")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("particle-steps/sec on 2D dam-break")
    assert d["steps"] == 1 and d["warmup"] == 3 and d["value"] > 0 and d["dtype"] == "f32" and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["unit"] == d["unit"] and "sample" in cb
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["scaling"] == "strong" and d["config"]["avg_level_sweeps"] >= 1 and d["config"]["particles_first"] > 0


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_weak_scaling_scene_columns():
    """The widened tanks of the round-1 weak-scaling runs (kept for tools/): N x the fluid of configs[1] as dam-break columns whose middles the equal-count slab faces
    (host mirror of dist.cu `rebalance`) cut through, about the same number of particles on every GPU."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import asph_b200 as A
    import bench
    one = A.scene_particle_count(bench.dam_break(A, bench.SPACING_C2))
    assert one == 999292
    for n in (2, 4, 8):
        sc = bench.dam_break(A, bench.SPACING_C2, n_gpus=n, kind="columns")
        assert len(sc.blocks) == n + 1 and abs(A.scene_particle_count(sc) / (n * one) - 1) < 2e-3
        assert A.scene_particle_count(bench.dam_break(A, bench.SPACING_C2, n_gpus=n, kind="wide")) >= n * one
        coarse = bench.dam_break(A, 8e-3, n_gpus=n, kind="columns")
        pos, _, _ = A.scene_particles(coarse)
        x = pos[:, 0].astype(np.float64)
        assert np.all(np.diff(x) >= 0)   # reference order is already sorted by x: contiguous index shares are x-slabs
        hist, _ = np.histogram(x, bins=8192, range=(x.min(), x.max()))
        faces = A.slab_bounds_from_histogram(hist, x.min(), x.max(), n)[1:-1]
        centres = [-n + 2.0 * k for k in range(1, n)]
        assert len(faces) == n - 1
        for f, c in zip(faces, centres):
            assert abs(f - c) < 0.05, (n, faces, centres)   # through the middle of a full column (0.7 wide)
        owned = np.histogram(x, bins=[-np.inf] + list(faces) + [np.inf])[0]
        assert owned.max() / owned.min() < 1.03, owned


def test_issue_roofline_arithmetic():
    """The second (issue-slot) reading of the Jacobi update pass next to the contract's HBM roofline: numbers of the round's
    measurement (36.5 us per launch, 18.4 M warp instructions, 148 SMs at 1965 MHz, 6456.8 GB/s)."""
    sys.path.insert(0, ROOT)
    import bench
    r = bench.issue_roofline(999292, 0.0365, 148, 1965.0, 18.40e6, 6456.8)
    assert abs(r["peak"] - 4 * 148 * 1.965) < 1e-6 and r["unit"] == "G warp-instructions/s"
    assert abs(r["frac"] - 0.4334) < 1e-3 and abs(r["instructions_per_particle"] - 589.2) < 0.1
    assert abs(r["hbm_frac_at_full_issue"] - 0.391) < 1e-3   # the HBM fraction is capped at 39 % by the instruction count
