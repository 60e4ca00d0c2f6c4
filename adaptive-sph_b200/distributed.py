"""Multi-GPU host side: one process per GPU (torchrun), slabs along x.

The reference is a single-process program; this is the host harness of the library's distributed handle
(`asph_create_distributed`, include/asph.h).  torch.distributed is plumbing only: rendezvous, broadcasting the NCCL
unique id, and gathering read-backs for tests.  All data-path communication (migration, ghost halos, the dt / Jacobi
statistics reductions) happens inside the library on its own CUDA stream.
"""
import ctypes as C
import os

import numpy as np

from .binding import FluidSimulation, load_library
from .scene import scene_boundary, scene_particle_count, scene_particles


def share_range(n_global, rank, world):
    """Contiguous share [lo, hi) of the reference particle order for `rank` (x-major lattices => already x-slabs)."""
    lo = (n_global * rank) // world
    hi = (n_global * (rank + 1)) // world
    return lo, hi


def slab_bounds_from_histogram(hist, x_min, x_max, world):
    """Host mirror of the library's slab-face rule (dist.cu `rebalance`): faces at equal particle counts, linear inside
    a histogram bin.  Returns world + 1 floats, first = -inf, last = +inf."""
    hist = np.asarray(hist, dtype=np.uint64)
    total = int(hist.sum())
    bins = len(hist)
    binw = np.float32((np.float32(x_max) - np.float32(x_min)) / np.float32(bins))
    out = [-np.inf]
    cum, b = 0, 0
    for r in range(1, world):
        target = (total * r) // world
        while b < bins and cum + int(hist[b]) < target:
            cum += int(hist[b]); b += 1
        frac = np.float32((target - cum) / int(hist[b])) if b < bins and hist[b] > 0 else np.float32(0)
        out.append(float(np.float32(x_min) + (np.float32(b) + frac) * binw))
    out.append(np.inf)
    return out


def owner_of(x, bounds):
    """Rank owning coordinate x: bounds[r] <= x < bounds[r + 1]."""
    return int(np.searchsorted(np.asarray(bounds[1:-1], dtype=np.float64), x, side="right"))


def broadcast_unique_id(lib, rank, src=0):
    """128-byte NCCL unique id from rank `src` to everyone (any torch.distributed backend)."""
    import torch.distributed as dist
    payload = [None]
    if rank == src:
        buf = (C.c_uint8 * 128)()
        rc = lib.asph_comm_unique_id(buf)
        if rc != 0:
            raise RuntimeError(f"asph_comm_unique_id failed ({rc}): NCCL not loadable?")
        payload[0] = bytes(buf)
    dist.broadcast_object_list(payload, src=src)
    return payload[0]


def gather_by_global_index(local_values, local_gidx, n_global, group=None):
    """Assemble a per-particle array in reference (global) order from every rank's owned part.  Works on any backend
    through all_gather_object (test / read-back path, not a hot path)."""
    import torch.distributed as dist
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, (np.asarray(local_gidx), np.asarray(local_values)), group=group)
    first = parts[0][1]
    out = np.zeros((n_global,) + first.shape[1:], dtype=first.dtype)
    seen = np.zeros(n_global, dtype=np.int32)
    for gidx, vals in parts:
        out[gidx] = vals
        seen[gidx] += 1
    if not np.all(seen == 1):
        raise RuntimeError(f"ownership is not a partition: {int((seen == 0).sum())} particles unowned, "
                           f"{int((seen > 1).sum())} owned more than once")
    return out


class DistributedFluidSimulation(FluidSimulation):
    """FluidSimulation over `world` GPUs.  Every rank constructs it with ITS share of the particles and their global
    (reference-order) indices; the library migrates particles to their owner slab at the first step."""

    def __init__(self, params, pos, vel, mass, global_index, n_global, boundary=None, counters_enabled=False,
                 capacity=0, lib=None, rank=None, world=None, device=None, split_patterns=None):
        import torch.distributed as dist
        lib = lib if lib is not None else load_library()
        rank = dist.get_rank() if rank is None else rank
        world = dist.get_world_size() if world is None else world
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", rank))
        self.rank, self.world, self.n_global = rank, world, int(n_global)
        nccl_id = broadcast_unique_id(lib, rank)
        super().__init__(params, pos, vel, mass, boundary, split_patterns, counters_enabled, capacity, lib,
                         distributed=dict(global_index=global_index, nccl_id=nccl_id, n_global=n_global, rank=rank,
                                          n_ranks=world, device=device))

    @classmethod
    def from_scene(cls, params, scene, **kw):
        import torch.distributed as dist
        rank = kw.get("rank", dist.get_rank())
        world = kw.get("world", dist.get_world_size())
        n_global = scene_particle_count(scene)
        lo, hi = share_range(n_global, rank, world)
        pos, vel, mass = scene_particles(scene, (lo, hi))
        gidx = np.arange(lo, hi, dtype=np.uint32)
        return cls(params, pos, vel, mass, gidx, n_global, scene_boundary(scene, params["init_boundary_handler"]), **kw)

    def num_global_particles(self):
        """Particles of the whole fluid now (resampling changes the count): the sum of what the ranks own."""
        import torch.distributed as dist
        counts = [None] * self.world
        dist.all_gather_object(counts, self.num_fluid_particles())
        return int(sum(counts))

    def gather_field(self, name):
        """Field of ALL particles in reference order, assembled on every rank."""
        return gather_by_global_index(self.get_field(name), self.global_index(), self.num_global_particles())
