"""The distributed handle with ONE rank (no neighbours, NCCL communicator of size 1) against the plain handle: exercises
the ownership split, the global-index cell sort, the read-back map and every collective call site on a single GPU."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_one_rank_distributed_equals_plain(asph, cuda_lib, default_params):
    params = default_params.replace(merging=False, sharing=False, splitting=False, level_estimation_method="None")
    scene = asph.SceneConfig.dam_break(0.008, pos=(-0.9, -0.6), size=(0.8, 0.4))
    pos, vel, mass = asph.scene_particles(scene)
    vel = (np.random.default_rng(3).standard_normal(vel.shape) * 0.05).astype(np.float32)  # the solver iterates from step 1
    b = asph.scene_boundary(scene, "AnalyticOverestimate")
    buf = (C.c_uint8 * 128)()
    assert cuda_lib.asph_comm_unique_id(buf) == 0
    n = len(mass)
    perm = np.random.default_rng(1).permutation(n).astype(np.uint32)  # hand the particles over in scrambled order
    d = asph.FluidSimulation(params, pos[perm], vel[perm], mass[perm], b, lib=cuda_lib,
                             distributed=dict(global_index=perm, nccl_id=bytes(buf), n_global=n, rank=0, n_ranks=1, device=0))
    s = asph.FluidSimulation(params, pos, vel, mass, b, lib=cuda_lib)
    most = 0
    for _ in range(4):  # few steps: an SPH compression amplifies the rounding differences of the two sort orders
        dt_d, dt_s = d.single_step(), s.single_step()
        assert dt_d == dt_s
        assert d.step_info()["density_sweeps"] == s.step_info()["density_sweeps"]
        assert d.step_info()["div_sweeps"] == s.step_info()["div_sweeps"]
        most = max(most, s.step_info()["density_sweeps"], s.step_info()["div_sweeps"])
    assert d.num_fluid_particles() == n
    gidx = d.global_index()
    assert np.array_equal(np.sort(gidx), np.arange(n, dtype=np.uint32))
    for name, scale, tol in (("position", 2.0, 1e-6), ("density", 1.0, 1e-4), ("mass", 1.0, 0.0)):
        a = np.empty_like(s.get_field(name)); a[gidx] = d.get_field(name)
        err = np.abs(a.astype(np.float64) - s.get_field(name)).max() / scale
        assert err <= tol, (name, err)
    assert most > 3
    d.close(); s.close()
