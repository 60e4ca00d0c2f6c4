"""VTK snapshots of the simulation state — the reference's on-disk format, and a restart path on top of it.

Mirrors platform/desktop/vtk_exporter.rs:
  * `write_vtk_file2(path, positions, data_ft, data_vec, data_u8, lines)` — vtk_exporter.rs:256-367: legacy VTK
    ("# vtk DataFile Version 4.2", title "SPH Particles 1.0", BINARY = big endian, DATASET POLYDATA): POINTS (2-D
    positions padded with z = 0, then the end points of the boundary lines), VERTICES (one per particle), LINES (one
    per boundary edge) and POINT_DATA with one SCALARS array per field — float fields with 1 component, vector
    fields with 3 (`DataArray::scalars(name, 3)`), flags as unsigned_char — each padded with zeros for the line points.
  * `write_vtk_file(path, sim, params)` — vtk_exporter.rs:81-167: the field list of the reference (density,
    density_error, density_error2, pressure, mass, aii, h, ppe_source_term; velocity, pressure_accel;
    flag_is_fluid_surface, flag_neighborhood_reduced; distances, lambda) and the polygon boundary's edges as lines.
    Fields a backend does not expose (the visualisation-only density_error arrays of the CUDA library) are omitted.
    The reference panics (`todo!()`) for plane boundaries; here those simply contribute no lines.
  * `VtkExporter` — vtk_exporter.rs:17-79, 248-253: numbered snapshots plus the `<basename>.vtk.series` JSON index
    ParaView reads.
The reference's writer is the `vtkio` crate (Cargo.lock: vtkio 0.6.3), which is not available here: the dataset layout
above follows its legacy writer, byte-level identity with it is unverified.

Added: `read_vtk_file` and `init_fluid_sim_from_vtk` — the persistent state of the step loop is exactly (x, v, m)
(SURVEY.md §8a), all three are in a snapshot as exact fp32, so a snapshot is a checkpoint: a simulation restarted from
it continues bit for bit like the uninterrupted one (tests/test_vtk.py).  The reference has no checkpoint / resume.
"""
import os

import numpy as np

VTK_HEADER = "# vtk DataFile Version 4.2\nSPH Particles 1.0\nBINARY\nDATASET POLYDATA\n"
_BND_PLANES, _BND_POLYGON = 1, 2


def _vec3(a):
    a = np.asarray(a, dtype=np.float32).reshape(-1, 2)
    out = np.zeros((len(a), 3), dtype=np.float32)
    out[:, :2] = a
    return out


def write_vtk_file2(path, positions, data_ft, data_vec, data_u8, lines):
    """vtk_exporter.rs:256-367.  positions [n, 2]; data_* = lists of (name, array); lines = list of ((x, y), (x, y))."""
    positions = np.asarray(positions, dtype=np.float32).reshape(-1, 2)
    n = len(positions)
    nl = len(lines)
    pts = [positions]
    for a, b in lines:
        pts.append(np.asarray([a, b], dtype=np.float32).reshape(2, 2))
    pts = _vec3(np.concatenate(pts)) if nl else _vec3(positions)
    with open(path, "wb") as f:
        f.write(VTK_HEADER.encode())
        f.write(f"POINTS {len(pts)} float\n".encode())
        f.write(pts.astype(">f4").tobytes()); f.write(b"\n")
        verts = np.empty((n, 2), dtype=">i4"); verts[:, 0] = 1; verts[:, 1] = np.arange(n)
        f.write(f"VERTICES {n} {2 * n}\n".encode())
        f.write(verts.tobytes()); f.write(b"\n")
        if nl:
            li = np.empty((nl, 3), dtype=">i4"); li[:, 0] = 2
            li[:, 1] = n + 2 * np.arange(nl); li[:, 2] = n + 2 * np.arange(nl) + 1
            f.write(f"LINES {nl} {3 * nl}\n".encode())
            f.write(li.tobytes()); f.write(b"\n")
        f.write(f"POINT_DATA {len(pts)}\n".encode())
        for name, arr in data_ft:
            arr = np.concatenate([np.asarray(arr, dtype=np.float32).reshape(-1), np.zeros(2 * nl, np.float32)])
            assert len(arr) == len(pts), name
            f.write(f"SCALARS {name} float 1\nLOOKUP_TABLE default\n".encode())
            f.write(arr.astype(">f4").tobytes()); f.write(b"\n")
        for name, arr in data_vec:
            arr = np.concatenate([_vec3(arr), np.zeros((2 * nl, 3), np.float32)])
            assert len(arr) == len(pts), name
            f.write(f"SCALARS {name} float 3\nLOOKUP_TABLE default\n".encode())
            f.write(arr.astype(">f4").tobytes()); f.write(b"\n")
        for name, arr in data_u8:
            arr = np.concatenate([np.asarray(arr, dtype=np.uint8).reshape(-1), np.zeros(2 * nl, np.uint8)])
            assert len(arr) == len(pts), name
            f.write(f"SCALARS {name} unsigned_char 1\nLOOKUP_TABLE default\n".encode())
            f.write(arr.tobytes()); f.write(b"\n")


def boundary_lines(boundary):
    """Sdf2D::draw_lines (sdf/sdf2d.rs:166-179): the polygon's edges; plane boundaries have none."""
    if boundary is None or boundary.kind != _BND_POLYGON:
        return []
    pts = [(float(boundary.poly[k][0]), float(boundary.poly[k][1])) for k in range(boundary.n_poly)]
    return [(pts[k], pts[(k + 1) % len(pts)]) for k in range(len(pts))]


def distance_to_boundary(boundary, pos):
    """BoundaryWinchenbach2020::distance_to_boundary (boundary_winchenbach2020.rs:308-325): min over the SDFs of probe(x).
    Host-side numpy (a visualisation field of the snapshot, not part of the step)."""
    pos = np.asarray(pos, dtype=np.float32).reshape(-1, 2)
    if boundary is None or boundary.kind not in (_BND_PLANES, _BND_POLYGON):
        return np.full(len(pos), np.inf, dtype=np.float32)
    if boundary.kind == _BND_PLANES:
        d = np.full(len(pos), np.inf, dtype=np.float32)
        for s in range(boundary.n_planes):
            nx, ny, dl = (np.float32(boundary.planes[s][k]) for k in range(3))
            d = np.minimum(d, nx * pos[:, 0] + ny * pos[:, 1] + dl)
        return d
    pts = np.array([[boundary.poly[k][0], boundary.poly[k][1]] for k in range(boundary.n_poly)], dtype=np.float32)
    nxt = np.roll(pts, -1, axis=0)
    edge = nxt - pts
    elen2 = (edge ** 2).sum(1)
    edir = edge / np.sqrt(elen2)[:, None]
    prev_dir = np.roll(edir, 1, axis=0)
    pn = np.stack([-prev_dir[:, 1] - edir[:, 1], prev_dir[:, 0] + edir[:, 0]], axis=1)
    best = np.full(len(pos), np.inf, dtype=np.float32)
    out = np.zeros(len(pos), dtype=np.float32)
    for k in range(len(pts)):  # find_min_dist_object / to_dist_and_dir, sdf2d.rs:73-141
        pd = pos - pts[k]
        proj = pd @ edir[k]
        dl = pd[:, 0] * -edir[k, 1] + pd[:, 1] * edir[k, 0]
        on = (proj > 0) & (proj * proj < elen2[k]) & (dl * dl < best)
        out = np.where(on, dl, out); best = np.where(on, dl * dl, best)
        c = (pd ** 2).sum(1)
        corner = c < best
        sign = np.where(pd @ pn[k] >= 0, 1.0, -1.0).astype(np.float32)
        out = np.where(corner, np.sqrt(c) * sign, out); best = np.where(corner, c, best)
    return out.astype(np.float32)


def write_vtk_file(path, sim, params=None, positions=None):
    """vtk_exporter.rs:81-167 for a FluidSimulation (binding.py).  Call it where the reference's exporter does: after
    `single_step_without_adaptivity`, before `single_step_adaptivity` (platform/desktop/animation/mod.rs:138-273) —
    the per-step fields describe the particle set of the physics step.  `positions` replaces the point coordinates
    (the batch exporter's interpolated frame positions, animation/mod.rs:193-210)."""
    from .binding import AsphError

    def opt(name):
        try:
            return sim.get_field(name)
        except AsphError:
            return None

    pos = sim.get_field("position") if positions is None else np.asarray(positions, np.float32)
    n = len(pos)
    data_ft, data_vec, data_u8 = [], [], []
    for vtk_name, field in (("density", "density"), ("density_error", "density_error"), ("density_error2", None),
                            ("pressure", "pressure"), ("mass", "mass"), ("aii", "aii"), ("h", "h"),
                            ("ppe_source_term", "ppe_source_term")):
        arr = opt(field) if field else None
        if arr is not None:
            data_ft.append((vtk_name, arr))
    data_vec.append(("velocity", sim.get_field("velocity")))
    pa = opt("pressure_accel")
    if pa is not None:
        data_vec.append(("pressure_accel", pa))
    fl = opt("flag_is_fluid_surface")
    data_u8.append(("flag_is_fluid_surface", fl if fl is not None else np.zeros(n, np.uint8)))
    data_u8.append(("flag_neighborhood_reduced", np.zeros(n, np.uint8)))  # only set by constrain_neighborhood_count (not built)
    b = getattr(sim, "_boundary", None)
    if b is not None and b.kind in (_BND_PLANES, _BND_POLYGON):
        data_ft.append(("distances", distance_to_boundary(b, pos)))
        lam = opt("lambda_sum")
        if lam is not None:
            data_ft.append(("lambda", lam))
    write_vtk_file2(path, pos, data_ft, data_vec, data_u8, boundary_lines(b))


class VtkExporter:
    """vtk_exporter.rs:17-79: `<folder>/<basename>-00001.vtk`, ... and `<folder>/<basename>.vtk.series`."""

    def __init__(self, folder, basename="sph"):
        self.folder, self.basename, self.snapshot_number = folder, basename, 1
        os.makedirs(folder, exist_ok=True)
        self.series_path = os.path.join(folder, basename + ".vtk.series")
        self._entries = []
        self._flush()

    def _flush(self):
        with open(self.series_path, "w") as f:
            f.write('{\n"file-series-version": "1.0",\n"files": [')
            f.write(",".join('\n{{ "name": "{}", "time": {} }}'.format(n, t) for n, t in self._entries))
            f.write("\n]\n}")

    def add_snapshot(self, time, sim, params=None):
        name = "{}-{:05d}.vtk".format(self.basename, self.snapshot_number)
        write_vtk_file(os.path.join(self.folder, name), sim, params)
        self._entries.append((name, str(np.float32(time))))  # the reference formats an f32 with `{}`
        self.snapshot_number += 1
        self._flush()
        return os.path.join(self.folder, name)


def read_vtk_file(path):
    """Reads what `write_vtk_file2` writes.  Returns {"position": [n, 2], "lines": [m, 2, 2], "<array name>": ...} with
    the per-particle arrays cut back to the n particles (the line end points carry dummy zeros)."""
    with open(path, "rb") as f:
        raw = f.read()
    at = 0

    def line():
        nonlocal at
        e = raw.index(b"\n", at)
        s = raw[at:e].decode()
        at = e + 1
        return s

    def blob(count, dtype):
        nonlocal at
        nbytes = count * np.dtype(dtype).itemsize
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=at)
        at += nbytes
        if raw[at:at + 1] == b"\n":
            at += 1
        return a

    if not line().startswith("# vtk DataFile Version"):
        raise ValueError("not a legacy VTK file")
    line()
    if line().strip() != "BINARY" or line().strip() != "DATASET POLYDATA":
        raise ValueError("expected BINARY / DATASET POLYDATA")
    out, n_pts, n = {}, 0, None
    types = {"float": ">f4", "unsigned_char": "u1", "double": ">f8", "int": ">i4"}
    while at < len(raw):
        head = line().split()
        if not head:
            continue
        if head[0] == "POINTS":
            n_pts = int(head[1])
            pts = blob(3 * n_pts, types[head[2]]).reshape(n_pts, 3).astype(np.float32)
        elif head[0] == "VERTICES":
            n = int(head[1])
            blob(int(head[2]), ">i4")
        elif head[0] == "LINES":
            m = int(head[1])
            li = blob(int(head[2]), ">i4").reshape(m, 3)
            out["lines"] = np.stack([pts[li[:, 1], :2], pts[li[:, 2], :2]], axis=1)
        elif head[0] == "POINT_DATA":
            assert int(head[1]) == n_pts
        elif head[0] == "SCALARS":
            name, typ, comps = head[1], head[2], int(head[3]) if len(head) > 3 else 1
            line()  # LOOKUP_TABLE default
            a = blob(comps * n_pts, types[typ])
            a = a.reshape(n_pts, comps)[:n] if comps > 1 else a[:n]
            if comps == 3:
                a = a[:, :2]
            out[name] = np.ascontiguousarray(a.astype(np.float32) if typ in ("float", "double") else a)
        else:
            raise ValueError("unexpected section " + head[0])
    if n is None:
        n = n_pts
    out["position"] = np.ascontiguousarray(pts[:n, :2])
    out.setdefault("lines", np.zeros((0, 2, 2), np.float32))
    return out


def init_fluid_sim_from_vtk(params, scene, path, split_patterns=None, counters_enabled=False, lib=None, capacity=0):
    """Restart: like `init_fluid_sim` (simulation.rs:3074), but the particles come from a snapshot (position, velocity,
    mass — the whole persistent state of the step loop) instead of the scene's blocks; the scene gives the boundary."""
    from .binding import FluidSimulation
    from .scene import scene_boundary
    d = read_vtk_file(path)
    boundary = scene_boundary(scene, params["init_boundary_handler"])
    return FluidSimulation(params, d["position"], d["velocity"], d["mass"], boundary, split_patterns, counters_enabled, capacity, lib)
