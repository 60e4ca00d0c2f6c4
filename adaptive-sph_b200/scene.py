"""Scene loading: SceneConfig YAML -> particle arrays + boundary description.

Mirrors (reference src/simulation/simulation.rs) `SceneConfig` :3052-3072, `add_fluid_block` :2915-2983
(fp32 lattice fill, x-major order), the boundary set-up of `init_fluid_sim` :3137-3213 and
`init_simulation_params` :3233-3256.  All arithmetic that decides counts / coordinates is done in
numpy float32 exactly as the Rust f32 code does it.
"""
import ctypes as C

import numpy as np
import yaml

f32 = np.float32

ASPH_MAX_PLANES = 8
ASPH_MAX_POLY_VERTS = 64
BND_NONE, BND_PLANES, BND_POLYGON = 0, 1, 2


class AsphBoundary(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n_planes", C.c_int32),
        ("planes", (C.c_float * 3) * ASPH_MAX_PLANES),
        ("n_poly", C.c_int32),
        ("poly", (C.c_float * 2) * ASPH_MAX_POLY_VERTS),
    ]


class SceneConfig:
    def __init__(self, mapping):
        b = mapping["boundary"]
        self.boundary_type = str(b["type"])
        self.width = f32(b["width"])
        self.height = f32(b["height"])
        self.blocks = []
        for blk in mapping["blocks"]:
            self.blocks.append(dict(
                pos=(f32(blk["pos"][0]), f32(blk["pos"][1])),
                size=(f32(blk["size"][0]), f32(blk["size"][1])),
                spacing=f32(blk["spacing"]),
                volume_fill_ratio=f32(blk["volume_fill_ratio"]),
                velocity=(f32(blk["velocity"][0]), f32(blk["velocity"][1])),
            ))

    @classmethod
    def from_yaml(cls, path):
        with open(path) as f:
            return cls(yaml.safe_load(f))

    @classmethod
    def dam_break(cls, spacing, pos=(-0.95, -0.9), size=(0.7, 1.8), width=2.0, height=2.0, fill=0.93):
        """`default-scene-web.yaml` geometry at another spacing (SURVEY.md §8 scale-ups)."""
        return cls({"boundary": {"type": "box", "width": width, "height": height},
                    "blocks": [{"pos": list(pos), "size": list(size), "spacing": spacing,
                                "volume_fill_ratio": fill, "velocity": [0, 0]}]})


def add_fluid_block(block):
    """simulation.rs:2915-2983.  Returns (pos[n,2], vel[n,2], mass[n]) float32, x-major order."""
    spacing = f32(block["spacing"])
    mn = np.array(block["pos"], dtype=f32)
    mx = np.array([block["pos"][0] + block["size"][0], block["pos"][1] + block["size"][1]], dtype=f32)
    particle_volume = f32(f32(spacing * spacing) * f32(block["volume_fill_ratio"]))
    particle_mass = f32(particle_volume * f32(1.0))  # INIT_REST_DENSITY
    box = (mx - mn).astype(f32)
    nx = int(np.floor(f32(box[0] / spacing)))
    ny = int(np.floor(f32(box[1] / spacing)))
    xs = (np.arange(nx, dtype=f32) * spacing + mn[0]).astype(f32)
    ys = (np.arange(ny, dtype=f32) * spacing + mn[1]).astype(f32)
    pos = np.empty((nx * ny, 2), dtype=f32)
    pos[:, 0] = np.repeat(xs, ny)
    pos[:, 1] = np.tile(ys, nx)
    vel = np.empty((nx * ny, 2), dtype=f32)
    vel[:, 0] = block["velocity"][0]
    vel[:, 1] = block["velocity"][1]
    mass = np.full(nx * ny, particle_mass, dtype=f32)
    return pos, vel, mass


def block_lattice(block):
    """(nx, ny, xs, ys, particle_mass) of add_fluid_block's lattice (simulation.rs:2957-2971)."""
    spacing = f32(block["spacing"])
    mn = np.array(block["pos"], dtype=f32)
    mx = np.array([block["pos"][0] + block["size"][0], block["pos"][1] + block["size"][1]], dtype=f32)
    particle_volume = f32(f32(spacing * spacing) * f32(block["volume_fill_ratio"]))
    particle_mass = f32(particle_volume * f32(1.0))
    box = (mx - mn).astype(f32)
    nx = int(np.floor(f32(box[0] / spacing)))
    ny = int(np.floor(f32(box[1] / spacing)))
    xs = (np.arange(nx, dtype=f32) * spacing + mn[0]).astype(f32)
    ys = (np.arange(ny, dtype=f32) * spacing + mn[1]).astype(f32)
    return nx, ny, xs, ys, particle_mass


def scene_particle_count(scene):
    return sum(block_lattice(b)[0] * block_lattice(b)[1] for b in scene.blocks)


def scene_particles(scene, index_range=None):
    """All particles of the scene in the reference's order (blocks in file order, x-major inside a block), or only
    those with reference index in [lo, hi) — a rank of a multi-GPU run generates just its own share."""
    if index_range is None:
        ps, vs, ms = [], [], []
        for blk in scene.blocks:
            p, v, m = add_fluid_block(blk)
            ps.append(p); vs.append(v); ms.append(m)
        if not ps:
            return np.zeros((0, 2), f32), np.zeros((0, 2), f32), np.zeros(0, f32)
        return np.concatenate(ps), np.concatenate(vs), np.concatenate(ms)
    lo, hi = int(index_range[0]), int(index_range[1])
    ps, vs, ms = [np.zeros((0, 2), f32)], [np.zeros((0, 2), f32)], [np.zeros(0, f32)]
    base = 0
    for blk in scene.blocks:
        nx, ny, xs, ys, pm = block_lattice(blk)
        a, b = max(lo, base), min(hi, base + nx * ny)
        if a < b:
            idx = np.arange(a - base, b - base, dtype=np.int64)
            p = np.empty((len(idx), 2), dtype=f32)
            p[:, 0] = xs[idx // ny]
            p[:, 1] = ys[idx % ny]
            v = np.empty_like(p)
            v[:, 0] = blk["velocity"][0]; v[:, 1] = blk["velocity"][1]
            ps.append(p); vs.append(v); ms.append(np.full(len(idx), pm, dtype=f32))
        base += nx * ny
    return np.concatenate(ps), np.concatenate(vs), np.concatenate(ms)


def scene_boundary(scene, init_boundary_handler):
    """Boundary handler set-up of init_fluid_sim (simulation.rs:3137-3213).

    AnalyticOverestimate -> the 4 planes of SdfPlane::new_boundary_box (sdf/sdf_plane.rs:13-20);
    AnalyticUnderestimate -> the polygon of Sdf2D::new_boundary_box (sdf/sdf2d.rs:153-164);
    NoBoundary -> none; Particles -> out of scope.
    """
    b = AsphBoundary()
    mn = np.array([f32(0) - f32(scene.width / f32(2)), f32(0) - f32(scene.height / f32(2))], dtype=f32)
    mx = np.array([f32(0) + f32(scene.width / f32(2)), f32(0) + f32(scene.height / f32(2))], dtype=f32)
    kind = str(init_boundary_handler)
    if kind == "AnalyticOverestimate":
        b.kind = BND_PLANES
        b.n_planes = 4
        planes = [(1.0, 0.0, -mn[0]), (-1.0, 0.0, mx[0]), (0.0, 1.0, -mn[1]), (0.0, -1.0, mx[1])]
        for k, (nx, ny, d) in enumerate(planes):
            b.planes[k][0], b.planes[k][1], b.planes[k][2] = float(nx), float(ny), float(d)
    elif kind == "AnalyticUnderestimate":
        b.kind = BND_POLYGON
        pts = [(mn[0], mn[1]), (mx[0], mn[1]), (mx[0], mx[1]), (mn[0], mx[1])]
        b.n_poly = len(pts)
        for k, (x, y) in enumerate(pts):
            b.poly[k][0], b.poly[k][1] = float(x), float(y)
    elif kind == "NoBoundary":
        b.kind = BND_NONE
    else:
        raise NotImplementedError("init_boundary_handler: Particles is out of scope (unusable in the adaptive build, "
                                  "particle_boundary_handler.rs:95-98)")
    return b


def init_simulation_params(params, scene):
    """simulation.rs:3233-3256, adaptive build: params.h is not used and is forced to 0."""
    return params.replace(h=0.0)
