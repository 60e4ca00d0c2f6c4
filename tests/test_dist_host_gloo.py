"""N > 1 host path on CPU: two `gloo` ranks under torchrun run tests/gloo_worker.py (no GPU, no kernels)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(world):
    port = 29900 + (os.getpid() % 90)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "gloo_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("GLOO_REPORT ")][-1]
    return json.loads(line[len("GLOO_REPORT "):])


@pytest.mark.parametrize("world", [2, 3])
def test_host_side_of_the_slab_decomposition(world):
    rep = _torchrun(world)
    assert rep["world"] == world
    for k in ("ok_share", "ok_id", "ok_owner", "ok_gather", "ok_detect"):
        assert rep[k], rep
    assert sum(rep["per_rank"]) == rep["n_global"]
    assert max(rep["per_rank"]) - min(rep["per_rank"]) <= 0.1 * rep["n_global"] / world, rep   # faces balance the slabs
    assert rep["bounds"] == sorted(rep["bounds"])
    assert rep["value"] == pytest.approx(rep["expected_value"])


def test_share_range_is_a_partition(asph):
    for n in (0, 1, 7, 1035, 999292):
        for world in (1, 2, 3, 8):
            edges = [asph.share_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_slab_faces_and_owner(asph):
    hist = np.zeros(16, np.uint64); hist[4:8] = 100            # all particles in x in [-0.5, 0)
    b = asph.slab_bounds_from_histogram(hist, -1.0, 1.0, 4)
    assert b[0] == -np.inf and b[-1] == np.inf and len(b) == 5
    assert np.allclose(b[1:-1], [-0.375, -0.25, -0.125], atol=1e-6)
    assert [asph.owner_of(x, b) for x in (-0.9, -0.375, -0.3, -0.2, 0.5)] == [0, 1, 1, 2, 3]
