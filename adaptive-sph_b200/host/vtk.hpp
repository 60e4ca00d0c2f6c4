// vtk.hpp — VTK snapshots of the simulation state in the layout of the reference's exporter
// (platform/desktop/vtk_exporter.rs:17-167, 256-367): legacy VTK 4.2, BINARY (big endian), DATASET
// POLYDATA with POINTS (z = 0; the end points of the boundary lines appended), VERTICES, LINES and one SCALARS array per
// field (float x 1, float x 3 for vectors, unsigned_char for flags), plus the `<basename>.vtk.series` index.  Same bytes as
// adaptive-sph_b200/vtk.py writes (tests/test_native_host.py reads both back).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "host.hpp"

namespace host {

namespace vtk_detail {
inline void put_be32(std::vector<unsigned char>& out, uint32_t v) {
  out.push_back(uint8_t(v >> 24)); out.push_back(uint8_t(v >> 16)); out.push_back(uint8_t(v >> 8)); out.push_back(uint8_t(v));
}
inline void put_f(std::vector<unsigned char>& out, float f) { uint32_t u; std::memcpy(&u, &f, 4); put_be32(out, u); }
inline void text(std::vector<unsigned char>& out, const std::string& s) { out.insert(out.end(), s.begin(), s.end()); }
}  // namespace vtk_detail

struct VtkField { std::string name; int comps; std::vector<float> f; std::vector<unsigned char> u8; bool is_u8; };
typedef std::pair<std::pair<float, float>, std::pair<float, float>> VtkLine;

// write_vtk_file2, vtk_exporter.rs:256-367
inline void write_vtk_file2(const std::string& path, const std::vector<float>& pos_xy, const std::vector<VtkField>& fields,
                            const std::vector<VtkLine>& lines) {
  using namespace vtk_detail;
  const size_t n = pos_xy.size() / 2, nl = lines.size(), npts = n + 2 * nl;
  std::vector<unsigned char> out;
  out.reserve(64 + npts * 12 + n * 8 + fields.size() * npts * 12);
  text(out, "# vtk DataFile Version 4.2\nSPH Particles 1.0\nBINARY\nDATASET POLYDATA\n");
  text(out, "POINTS " + std::to_string(npts) + " float\n");
  for (size_t i = 0; i < n; i++) { put_f(out, pos_xy[2 * i]); put_f(out, pos_xy[2 * i + 1]); put_f(out, 0.f); }
  for (auto& l : lines) {
    put_f(out, l.first.first); put_f(out, l.first.second); put_f(out, 0.f);
    put_f(out, l.second.first); put_f(out, l.second.second); put_f(out, 0.f);
  }
  text(out, "\n");
  text(out, "VERTICES " + std::to_string(n) + " " + std::to_string(2 * n) + "\n");
  for (size_t i = 0; i < n; i++) { put_be32(out, 1u); put_be32(out, uint32_t(i)); }
  text(out, "\n");
  if (nl) {
    text(out, "LINES " + std::to_string(nl) + " " + std::to_string(3 * nl) + "\n");
    for (size_t k = 0; k < nl; k++) { put_be32(out, 2u); put_be32(out, uint32_t(n + 2 * k)); put_be32(out, uint32_t(n + 2 * k + 1)); }
    text(out, "\n");
  }
  text(out, "POINT_DATA " + std::to_string(npts) + "\n");
  for (int pass = 0; pass < 3; pass++) {  // float x 1, then float x 3, then unsigned_char, as the reference orders them
    for (auto& f : fields) {
      const int kind = f.is_u8 ? 2 : (f.comps == 1 ? 0 : 1);
      if (kind != pass) continue;
      if (f.is_u8) {
        text(out, "SCALARS " + f.name + " unsigned_char 1\nLOOKUP_TABLE default\n");
        out.insert(out.end(), f.u8.begin(), f.u8.end());
        out.insert(out.end(), 2 * nl, 0);
      } else if (f.comps == 1) {
        text(out, "SCALARS " + f.name + " float 1\nLOOKUP_TABLE default\n");
        for (float v : f.f) put_f(out, v);
        for (size_t k = 0; k < 2 * nl; k++) put_f(out, 0.f);
      } else {
        text(out, "SCALARS " + f.name + " float 3\nLOOKUP_TABLE default\n");
        for (size_t i = 0; i < n; i++) { put_f(out, f.f[2 * i]); put_f(out, f.f[2 * i + 1]); put_f(out, 0.f); }
        for (size_t k = 0; k < 6 * nl; k++) put_f(out, 0.f);
      }
      text(out, "\n");
    }
  }
  FILE* fp = std::fopen(path.c_str(), "wb");
  if (!fp) throw std::runtime_error("cannot write " + path);
  std::fwrite(out.data(), 1, out.size(), fp);
  std::fclose(fp);
}

// Sdf2D::draw_lines (sdf/sdf2d.rs:166-179): the polygon's edges; plane boundaries have none
inline std::vector<VtkLine> boundary_lines(const asph_boundary& b) {
  std::vector<VtkLine> lines;
  if (b.kind != ASPH_BND_POLYGON) return lines;
  for (int k = 0; k < b.n_poly; k++) {
    const int k1 = (k + 1) % b.n_poly;
    lines.push_back({{b.poly[k][0], b.poly[k][1]}, {b.poly[k1][0], b.poly[k1][1]}});
  }
  return lines;
}

// BoundaryWinchenbach2020::distance_to_boundary (boundary_winchenbach2020.rs:308-325): min over the SDFs of probe(x);
// planes sdf/sdf_plane.rs:36-38, polygon sdf/sdf2d.rs:73-141 (nearest edge or vertex, positive on the fluid side)
inline std::vector<float> distance_to_boundary(const asph_boundary& b, const std::vector<float>& pos_xy) {
  const size_t n = pos_xy.size() / 2;
  std::vector<float> d(n, std::numeric_limits<float>::infinity());
  if (b.kind == ASPH_BND_PLANES) {
    for (size_t i = 0; i < n; i++)
      for (int s = 0; s < b.n_planes; s++) {
        float v = b.planes[s][0] * pos_xy[2 * i];
        const float w = b.planes[s][1] * pos_xy[2 * i + 1];
        v = v + w;
        v = v + b.planes[s][2];
        d[i] = std::min(d[i], v);
      }
  } else if (b.kind == ASPH_BND_POLYGON) {
    const int np = b.n_poly;
    std::vector<float> ex(np), ey(np), len2(np), pnx(np), pny(np);
    for (int k = 0; k < np; k++) {
      const int k1 = (k + 1) % np;
      const float lx = b.poly[k1][0] - b.poly[k][0], ly = b.poly[k1][1] - b.poly[k][1];
      len2[k] = lx * lx + ly * ly;
      const float len = std::sqrt(len2[k]);
      ex[k] = lx / len; ey[k] = ly / len;
    }
    for (int k = 0; k < np; k++) {
      const int a = k == 0 ? np - 1 : k - 1;
      pnx[k] = -ey[a] - ey[k]; pny[k] = ex[a] + ex[k];
    }
    for (size_t i = 0; i < n; i++) {
      float best = std::numeric_limits<float>::infinity(), out = 0.f;
      for (int k = 0; k < np; k++) {
        const float px = pos_xy[2 * i] - b.poly[k][0], py = pos_xy[2 * i + 1] - b.poly[k][1];
        const float proj = px * ex[k] + py * ey[k];
        const float dl = px * -ey[k] + py * ex[k];
        if (proj > 0.f && proj * proj < len2[k] && dl * dl < best) { out = dl; best = dl * dl; }
        const float c = px * px + py * py;
        if (c < best) { out = std::sqrt(c) * ((px * pnx[k] + py * pny[k]) >= 0.f ? 1.f : -1.f); best = c; }
      }
      d[i] = out;
    }
  }
  return d;
}

// write_vtk_file, vtk_exporter.rs:81-167: the reference's field list; fields a backend does not expose are omitted.
// `positions` replaces the point coordinates (interpolated video frames of the batch exporter).
inline void write_vtk_file(const std::string& path, FluidSimulation& sim, const asph_boundary& boundary,
                           const std::vector<float>* positions = nullptr) {
  const std::vector<float> pos = positions ? *positions : sim.field(ASPH_F_POSITION, 2);
  const size_t n = pos.size() / 2;
  std::vector<VtkField> fields;
  auto try_float = [&](const char* name, int id, int comps) {
    VtkField f{name, comps, std::vector<float>(n * size_t(comps)), {}, false};
    if (sim.lib.get_field(sim.sim, id, f.f.data(), f.f.size() * sizeof(float)) == ASPH_OK) fields.push_back(std::move(f));
  };
  try_float("density", ASPH_F_DENSITY, 1);
  try_float("density_error", ASPH_F_DENSITY_ERROR, 1);
  try_float("pressure", ASPH_F_PRESSURE, 1);
  try_float("mass", ASPH_F_MASS, 1);
  try_float("aii", ASPH_F_AII, 1);
  try_float("h", ASPH_F_H, 1);
  try_float("ppe_source_term", ASPH_F_SOURCE_TERM, 1);
  try_float("velocity", ASPH_F_VELOCITY, 2);
  try_float("pressure_accel", ASPH_F_PRESSURE_ACCEL, 2);
  {
    VtkField f{"flag_is_fluid_surface", 1, {}, std::vector<unsigned char>(n, 0), true};
    if (sim.lib.get_field(sim.sim, ASPH_F_FLAG_SURFACE, f.u8.data(), f.u8.size()) != ASPH_OK) std::fill(f.u8.begin(), f.u8.end(), 0);
    fields.push_back(std::move(f));
    fields.push_back(VtkField{"flag_neighborhood_reduced", 1, {}, std::vector<unsigned char>(n, 0), true});
  }
  if (boundary.kind == ASPH_BND_PLANES || boundary.kind == ASPH_BND_POLYGON) {
    fields.push_back(VtkField{"distances", 1, distance_to_boundary(boundary, pos), {}, false});
    try_float("lambda", ASPH_F_LAMBDA_SUM, 1);
  }
  write_vtk_file2(path, pos, fields, boundary_lines(boundary));
}

// Reads what write_vtk_file2 writes, as far as a restart needs it: the positions of the n particles (VERTICES count) and the
// float arrays, cut back to the particles.  The persistent state of the step loop is exactly (x, v, m), all
// three are in a snapshot as exact fp32, so a snapshot is a checkpoint (the reference has no checkpoint / resume).
struct VtkSnapshot {
  std::vector<float> position;                                   // [2n]
  std::vector<std::pair<std::string, std::vector<float>>> arrays;  // name -> n (scalars) or 2n (vectors) values
  const std::vector<float>& get(const std::string& name) const {
    for (auto& a : arrays) if (a.first == name) return a.second;
    throw std::runtime_error("snapshot has no array `" + name + "`");
  }
};
inline VtkSnapshot read_vtk_file(const std::string& path) {
  FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) throw std::runtime_error("cannot read " + path);
  std::vector<unsigned char> raw;
  unsigned char buf[1 << 16];
  for (size_t k; (k = std::fread(buf, 1, sizeof(buf), fp)) > 0;) raw.insert(raw.end(), buf, buf + k);
  std::fclose(fp);
  size_t at = 0;
  auto line = [&]() {
    size_t e = at;
    while (e < raw.size() && raw[e] != '\n') e++;
    std::string s(raw.begin() + long(at), raw.begin() + long(e));
    at = std::min(raw.size(), e + 1);
    return s;
  };
  auto words = [](const std::string& s) {
    std::vector<std::string> w;
    std::istringstream in(s);
    for (std::string t; in >> t;) w.push_back(t);
    return w;
  };
  auto be_float = [&](size_t off) {
    const uint32_t u = (uint32_t(raw[off]) << 24) | (uint32_t(raw[off + 1]) << 16) | (uint32_t(raw[off + 2]) << 8) | uint32_t(raw[off + 3]);
    float f;
    std::memcpy(&f, &u, 4);
    return f;
  };
  auto skip = [&](size_t nbytes) {
    if (at + nbytes > raw.size()) throw std::runtime_error("truncated VTK file " + path);
    at += nbytes;
    if (at < raw.size() && raw[at] == '\n') at++;
  };
  if (line().compare(0, 22, "# vtk DataFile Version") != 0) throw std::runtime_error("not a legacy VTK file: " + path);
  line();
  if (words(line()) != std::vector<std::string>{"BINARY"} || words(line()) != std::vector<std::string>{"DATASET", "POLYDATA"})
    throw std::runtime_error("expected BINARY / DATASET POLYDATA: " + path);
  VtkSnapshot snap;
  size_t n_pts = 0, n = size_t(-1), pts_at = 0;
  while (at < raw.size()) {
    const std::vector<std::string> head = words(line());
    if (head.empty()) continue;
    if (head[0] == "POINTS") {
      if (head.size() < 3 || head[2] != "float") throw std::runtime_error("POINTS: expected float");
      n_pts = size_t(std::stoull(head[1]));
      pts_at = at;
      skip(12 * n_pts);
    } else if (head[0] == "VERTICES") {
      n = size_t(std::stoull(head[1]));
      skip(4 * size_t(std::stoull(head[2])));
    } else if (head[0] == "LINES") {
      skip(4 * size_t(std::stoull(head[2])));
    } else if (head[0] == "POINT_DATA") {
      if (size_t(std::stoull(head[1])) != n_pts) throw std::runtime_error("POINT_DATA count differs from POINTS");
    } else if (head[0] == "SCALARS") {
      const int comps = head.size() > 3 ? std::stoi(head[3]) : 1;
      line();  // LOOKUP_TABLE default
      const size_t keep = std::min(n, n_pts);
      if (head[2] == "float") {
        std::vector<float> a;
        const int take = comps == 3 ? 2 : comps;
        a.reserve(keep * size_t(take));
        for (size_t i = 0; i < keep; i++)
          for (int c = 0; c < take; c++) a.push_back(be_float(at + 4 * (i * size_t(comps) + size_t(c))));
        snap.arrays.push_back({head[1], std::move(a)});
        skip(4 * n_pts * size_t(comps));
      } else if (head[2] == "unsigned_char") {
        skip(n_pts * size_t(comps));
      } else {
        throw std::runtime_error("unsupported array type " + head[2]);
      }
    } else {
      throw std::runtime_error("unexpected section " + head[0]);
    }
  }
  const size_t keep = std::min(n, n_pts);
  snap.position.reserve(2 * keep);
  for (size_t i = 0; i < keep; i++) { snap.position.push_back(be_float(pts_at + 12 * i)); snap.position.push_back(be_float(pts_at + 12 * i + 4)); }
  return snap;
}

// VtkExporter, vtk_exporter.rs:17-79: `<folder>/<basename>-00001.vtk`, ... and `<folder>/<basename>.vtk.series`
struct VtkExporter {
  std::string folder, basename;
  int snapshot_number = 1;
  std::vector<std::pair<std::string, std::string>> entries;
  VtkExporter(const std::string& f, const std::string& b) : folder(f), basename(b) { flush(); }
  void flush() const {
    FILE* fp = std::fopen((folder + "/" + basename + ".vtk.series").c_str(), "w");
    if (!fp) throw std::runtime_error("cannot write into " + folder);
    std::fputs("{\n\"file-series-version\": \"1.0\",\n\"files\": [", fp);
    for (size_t k = 0; k < entries.size(); k++)
      std::fprintf(fp, "%s\n{ \"name\": \"%s\", \"time\": %s }", k ? "," : "", entries[k].first.c_str(), entries[k].second.c_str());
    std::fputs("\n]\n}", fp);
    std::fclose(fp);
  }
  std::string add_snapshot(double time, FluidSimulation& sim, const asph_boundary& boundary) {
    char name[512], t[64];
    std::snprintf(name, sizeof(name), "%s-%05d.vtk", basename.c_str(), snapshot_number);
    write_vtk_file(folder + "/" + name, sim, boundary);
    std::snprintf(t, sizeof(t), "%.9g", double(float(time)));
    entries.push_back({name, t});
    snapshot_number++;
    flush();
    return folder + "/" + name;
  }
};

}  // namespace host
