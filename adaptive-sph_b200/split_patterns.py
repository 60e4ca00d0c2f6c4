"""split-patterns.yaml -> flat arrays for `asph_split_patterns` (include/asph.h).

Mirrors SplitPatterns (reference src/simulation/adaptivity/splitting.rs:84-120) and
load_split_patterns_from_file (simulation.rs:3000-3004): a YAML list whose entry k is the 1 -> (k+2) pattern
with fields mass_s, pos_s, h_s; only pos_s is used at run time.
"""
import ctypes as C
import os

import numpy as np
import yaml

DEFAULT_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "split-patterns.yaml")


class AsphSplitPatterns(C.Structure):
    _fields_ = [("max_children", C.c_int32), ("offset", C.POINTER(C.c_int32)), ("pos_xy", C.POINTER(C.c_float))]


class SplitPatterns:
    def __init__(self, patterns):
        # SplitPatterns::new asserts pattern i has i + 2 points (splitting.rs:102-108)
        for i, sp in enumerate(patterns):
            if len(sp["pos_s"]) != i + 2:
                raise ValueError(f"split pattern {i} has {len(sp['pos_s'])} points, expected {i + 2}")
        self.patterns = patterns
        self.max_children = len(patterns) + 1  # get_max_num_children splitting.rs:117-119
        offs, pos, o = [], [], 0
        for sp in patterns:
            offs.append(o)
            pos.extend(sp["pos_s"])
            o += len(sp["pos_s"])
        self.offset = np.asarray(offs, dtype=np.int32)
        self.pos = np.asarray(pos, dtype=np.float32).reshape(-1, 2)
        self.c = AsphSplitPatterns(self.max_children,
                                   self.offset.ctypes.data_as(C.POINTER(C.c_int32)),
                                   self.pos.ctypes.data_as(C.POINTER(C.c_float)))

    def get(self, num_children):
        return self.pos[self.offset[num_children - 2]: self.offset[num_children - 2] + num_children]


def load_split_patterns_from_file(path=None):
    path = path or DEFAULT_PATH
    cache = os.path.splitext(path)[0] + ".npz"
    with open(path) as f:
        # the C loader of PyYAML makes the 129 kB file load in milliseconds
        loader = getattr(yaml, "CSafeLoader", yaml.SafeLoader)
        data = yaml.load(f, Loader=loader)
    del cache
    return SplitPatterns(data)
