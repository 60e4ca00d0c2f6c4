"""world_size-2 `gloo` worker (CPU): the host-side logic of the multi-GPU path — share ranges, per-rank scene
generation, slab faces from the global x-histogram, ownership, NCCL-id broadcast plumbing, reference-order gather and
the bench's sum-of-work / max-of-time reduction.  No device, no compute: the CUDA library is not called."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeIdLib:
    """Stands in for libasph_b200's asph_comm_unique_id (which needs NCCL + a GPU): fills a recognisable 128-byte id."""
    def asph_comm_unique_id(self, buf):
        for k in range(128):
            buf[k] = (k * 7 + 3) & 0xFF
        return 0


def main():
    import torch
    import torch.distributed as dist
    import asph_b200 as A

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    scene = A.SceneConfig.dam_break(0.02, pos=(-0.9, -0.6), size=(1.2, 0.5))
    n_global = A.scene_particle_count(scene)
    lo, hi = A.share_range(n_global, rank, world)
    pos, vel, mass = A.scene_particles(scene, (lo, hi))
    full_pos, full_vel, full_mass = A.scene_particles(scene)
    ok_share = np.array_equal(pos, full_pos[lo:hi]) and np.array_equal(mass, full_mass[lo:hi])

    # NCCL unique id: created on rank 0 only, identical bytes everywhere
    nid = A.broadcast_unique_id(FakeIdLib(), rank)
    ok_id = len(nid) == 128 and nid == bytes((k * 7 + 3) & 0xFF for k in range(128))

    # slab faces from the summed histogram (what dist.cu's rebalance does with ncclAllReduce)
    bins, x_min, x_max = 64, -1.0, 1.0
    h = np.histogram(pos[:, 0], bins=bins, range=(x_min, x_max))[0].astype(np.int64)
    ht = torch.from_numpy(h)
    dist.all_reduce(ht)
    bounds = A.slab_bounds_from_histogram(ht.numpy(), x_min, x_max, world)
    owners_full = np.array([A.owner_of(x, bounds) for x in full_pos[:, 0]])
    per_rank = np.bincount(owners_full, minlength=world)

    # migration: every rank hands each of its particles to the owner of its x (all_to_all of index lists)
    own = np.array([A.owner_of(x, bounds) for x in pos[:, 0]], dtype=np.int64)
    gidx = np.arange(lo, hi, dtype=np.uint32)
    outbox = [(gidx[own == r], pos[own == r]) for r in range(world)]
    everything = [None] * world
    dist.all_gather_object(everything, outbox)
    mine_idx = np.concatenate([everything[src][rank][0] for src in range(world)])
    mine_pos = np.concatenate([everything[src][rank][1] for src in range(world)])
    ok_owner = all(A.owner_of(x, bounds) == rank for x in mine_pos[:, 0]) and len(mine_idx) == per_rank[rank]

    # read-back in reference order from the migrated ownership
    gathered = A.gather_by_global_index(mine_pos, mine_idx, n_global)
    ok_gather = np.array_equal(gathered, full_pos)

    # a non-partition must be detected
    try:
        A.gather_by_global_index(mine_pos[:-1] if rank == 0 else mine_pos, mine_idx[:-1] if rank == 0 else mine_idx, n_global)
        ok_detect = False
    except RuntimeError:
        ok_detect = True

    # bench reduction: value = sum over ranks of particle-steps / max over ranks of time
    work = torch.tensor([float(len(mine_idx) * 10)], dtype=torch.float64)
    t_ms = torch.tensor([5.0 + rank], dtype=torch.float64)
    dist.all_reduce(work, op=dist.ReduceOp.SUM)
    dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    value = float(work.item()) / (float(t_ms.item()) * 1e-3)

    if rank == 0:
        print("GLOO_REPORT " + json.dumps({
            "world": world, "n_global": int(n_global), "ok_share": bool(ok_share), "ok_id": bool(ok_id), "ok_owner": bool(ok_owner),
            "ok_gather": bool(ok_gather), "ok_detect": bool(ok_detect), "per_rank": per_rank.tolist(),
            "bounds": [float(b) for b in bounds[1:-1]], "value": value,
            "expected_value": n_global * 10 / ((5.0 + world - 1) * 1e-3)}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
