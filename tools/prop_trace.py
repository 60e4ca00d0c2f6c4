"""Development probe: per-sweep cycle counts of k_propagate (library built with `make EXTRA_level=-DASPH_PROP_TRACE`).

  python tools/prop_trace.py            # the bench workload after its pre-roll, one step traced
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asph_b200 as A
import bench


def main():
    lib = A.load_library()
    base = bench.adaptive_params(A, bench.SPACING)
    scene = bench.dam_break(A, bench.SPACING)
    base = A.init_simulation_params(base, scene)
    sim = A.init_fluid_sim(base, scene, A.load_split_patterns_from_file(), counters_enabled=True, lib=lib)
    bench.preroll_adaptive(sim, A, base, bench.SPACING)
    for _ in range(3):
        sim.single_step(base)
    buf = np.zeros((12, 512), dtype=np.uint64)
    ptr = buf.ctypes.data_as(C.POINTER(C.c_ulonglong))
    lib.asph_debug_prop_trace(ptr, 1)
    sim.single_step(base)
    lib.asph_debug_prop_trace(ptr, 0)
    nb = 296
    sw = int(sim.step_info()["level_sweeps"])
    print("sweeps", sw, "warps", int(buf[5, 1]))
    print(" t   front  push_max  push_mean  blk0_flush  blk0_total  (cycles)")
    for t in range(1, sw + 1):
        if t <= 12 or t % 10 == 0:
            print(f"{t:3d} {int(buf[0, t]):7d} {int(buf[1, t]):9d} {int(buf[2, t]) / max(1, int(buf[5, t])):10.0f} {int(buf[3, t]):10d} {int(buf[4, t]):10d}")
    r = slice(1, sw + 1)
    print("mean front", buf[0, r].mean(), "push_max", buf[1, r].mean(), "push_mean", (buf[2, r] / np.maximum(buf[5, r], 1)).mean(),
          "blk0 before sync", buf[3, r].mean(), "blk0 after sync", buf[4, r].mean())
    sim.close()


if __name__ == "__main__":
    main()
