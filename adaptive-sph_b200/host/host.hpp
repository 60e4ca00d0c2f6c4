// host.hpp — C++ host side above the C ABI of include/asph.h: what the reference's Rust front end does around
// FluidSimulation (no Rust toolchain exists in this image, so the host that a maintainer would write in Rust — see
// INTEGRATION.md — is C++ here; the Python package next to it is the test harness over the same ABI).
//
//   SimulationParams  <- YAML   simulation_parameters.rs:25-213 (field names == YAML keys; serde: every non-Option field
//                               is mandatory, unknown variants are errors), `-c` overwrite merge main_loop.rs:113-126
//   SceneConfig       <- YAML   simulation.rs:3052-3072; add_fluid_block :2915-2983 (fp32 lattice fill, x-major);
//                               boundary set-up of init_fluid_sim :3137-3213; init_simulation_params :3233-3256
//   SplitPatterns     <- YAML   adaptivity/splitting.rs:84-120, load_split_patterns_from_file simulation.rs:3000
//   Library                     dlopen of a library exporting include/asph.h (default: ../csrc/libasph_b200.so — the CUDA
//                               library, which has no CPU fallback)
//   Statistics                  ValueCounters + write_statistics simulation.rs:137-157, 3279-3359
//
// All arithmetic that decides particle counts / coordinates is float, operation by operation as the Rust f32 code does it
// (compile with -ffp-contract=off).
#pragma once
#include <dlfcn.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/asph.h"
#include "yaml_lite.hpp"

namespace host {

using yaml_lite::Node;

// ------------------------------------------------------------------------------------------------ SimulationParams
struct EnumSpec { const char* field; std::vector<const char*> variants; };
inline const std::vector<EnumSpec>& enum_specs() {
  static const std::vector<EnumSpec> specs = {
      {"viscosity_type", {"WCSPH", "ApproxLaplace", "XSPH"}},
      {"level_estimation_method", {"None", "CenterDiff", "EmptyAngle"}},
      {"neighborhood_search_algorithm", {"Grid", "RStar"}},
      {"init_boundary_handler", {"Particles", "AnalyticUnderestimate", "AnalyticOverestimate", "NoBoundary"}},
      {"support_length_estimation", {"FromDistribution", "FromDistributionClamped1", "FromDistributionClamped2", "FromDistribution2", "FromMass"}},
      {"pressure_solver_method", {"IISPH", "IISPH2", "HybridDFSPH", "OnlyDivergence"}},
      {"hybrid_dfsph_density_source_term", {"DensityAndDivergence", "OnlyDensity"}},
      {"boundary_penalty_term", {"None", "Linear", "Quadratic1", "Quadratic2"}},
      {"sizing_function", {"Radius2", "Radius", "Mass"}},
      {"operator_discretization", {"ConsistentSimpleGradient", "ConsistentSymmetricGradient", "Winchenbach2020"}},
  };
  return specs;
}
inline int enum_index(const std::string& field, const Node& v) {
  for (auto& s : enum_specs()) {
    if (field != s.field) continue;
    const std::string name = v.is_null() ? "None" : v.as_string();  // the variant called None
    for (size_t k = 0; k < s.variants.size(); k++) if (name == s.variants[k]) return int(k);
    throw std::runtime_error(field + ": unknown variant `" + name + "`");
  }
  throw std::runtime_error("not an enum field: " + field);
}

struct ParamField { const char* name; char type; size_t offset; };  // type: d double, b bool (int32), i int32, l int64, e enum
#define ASPH_PF(name, type) {#name, type, offsetof(asph_params, name)}
inline const std::vector<ParamField>& param_fields() {
  static const std::vector<ParamField> f = {
      ASPH_PF(rest_density, 'd'), ASPH_PF(cfl_factor, 'd'), ASPH_PF(max_dt, 'd'), ASPH_PF(h, 'd'), ASPH_PF(use_iisph, 'b'),
      ASPH_PF(viscosity, 'd'), ASPH_PF(viscosity_type, 'e'), ASPH_PF(gravity, 'd'), ASPH_PF(check_aii, 'b'),
      ASPH_PF(level_estimation_method, 'e'), ASPH_PF(maximum_range, 'd'), ASPH_PF(jacobi_omega, 'd'), ASPH_PF(eos_stiffness, 'd'),
      ASPH_PF(eos_power, 'i'), ASPH_PF(neighborhood_search_algorithm, 'e'), ASPH_PF(init_boundary_handler, 'e'),
      ASPH_PF(support_length_estimation, 'e'), ASPH_PF(sdf_gradient_eps, 'd'), ASPH_PF(fail_on_missing_split_pattern, 'b'),
      ASPH_PF(constrain_neighborhood_count, 'b'), ASPH_PF(particle_radius_fine, 'd'), ASPH_PF(particle_radius_base, 'd'),
      ASPH_PF(maximum_surface_distance, 'd'), ASPH_PF(minimum_share_partners, 'i'), ASPH_PF(minimum_merge_partners, 'i'),
      ASPH_PF(merging, 'b'), ASPH_PF(sharing, 'b'), ASPH_PF(splitting, 'b'), ASPH_PF(max_mass_transfer_sharing, 'd'),
      ASPH_PF(max_mass_transfer_merging, 'd'), ASPH_PF(max_share_distance, 'd'), ASPH_PF(max_merge_distance, 'd'),
      ASPH_PF(allow_merge_with_optimal_particle, 'b'), ASPH_PF(allow_share_with_optimal_particle, 'b'),
      ASPH_PF(allow_share_with_too_small_particle, 'b'), ASPH_PF(allow_merge_on_size_difference, 'b'),
      ASPH_PF(boundary_is_fluid_surface, 'b'), ASPH_PF(use_extended_range_for_level_estimation, 'b'),
      ASPH_PF(pressure_solver_method, 'e'), ASPH_PF(iisph_max_avg_density_error, 'd'), ASPH_PF(hybrid_dfsph_factor, 'd'),
      ASPH_PF(hybrid_dfsph_max_avg_density_error, 'd'), ASPH_PF(hybrid_dfsph_max_avg_divergence_error, 'd'),
      ASPH_PF(hybrid_dfsph_density_source_term, 'e'), ASPH_PF(hybrid_dfsph_non_pressure_accel_before_divergence_free, 'b'),
      ASPH_PF(check_neighborhood, 'b'), ASPH_PF(boundary_penalty_term, 'e'), ASPH_PF(sizing_function, 'e'),
      ASPH_PF(level_estimation_after_advection, 'b'), ASPH_PF(level_estimation_range, 'd'), ASPH_PF(operator_discretization, 'e'),
      ASPH_PF(max_iters, 'l'),
  };
  return f;
}
#undef ASPH_PF
inline bool optional_param(const std::string& k) {
  return k == "pull_fluid_to" || k == "fill_stash_with" || k == "operator_discretization_for_diagonal";
}

// `-c` semantics: every key of the overwrite mapping must already exist (main_loop.rs:119-124)
inline void merge_overwrite(Node& mapping, const Node& over) {
  if (over.is_null()) return;
  if (over.kind != Node::Map) throw std::runtime_error("overwrite file: expected a mapping");
  for (auto& kv : over.map) {
    Node* dst = mapping.find(kv.first);
    if (!dst) throw std::runtime_error("not able to find attribute " + kv.first);
    *dst = kv.second;
  }
}

inline asph_params params_from_yaml(const Node& m) {
  if (m.kind != Node::Map) throw std::runtime_error("failed to unpack SimulationParams: expected a mapping");
  asph_params p;
  std::memset(&p, 0, sizeof(p));
  std::string missing;
  for (auto& f : param_fields()) {
    const Node* v = m.find(f.name);
    if (!v) { missing += std::string(missing.empty() ? "" : ", ") + f.name; continue; }
    char* dst = reinterpret_cast<char*>(&p) + f.offset;
    try {
      switch (f.type) {
        case 'd': *reinterpret_cast<double*>(dst) = v->as_double(); break;
        case 'b': *reinterpret_cast<int32_t*>(dst) = v->as_bool() ? 1 : 0; break;
        case 'i': *reinterpret_cast<int32_t*>(dst) = int32_t(v->as_double()); break;
        case 'l': *reinterpret_cast<int64_t*>(dst) = int64_t(v->as_double()); break;
        case 'e': *reinterpret_cast<int32_t*>(dst) = enum_index(f.name, *v); break;
      }
    } catch (const std::exception& e) {
      throw std::runtime_error(std::string(f.name) + ": " + e.what());
    }
  }
  if (!missing.empty()) throw std::runtime_error("failed to unpack SimulationParams: missing field(s) " + missing);
  for (auto& kv : m.map) {
    bool known = optional_param(kv.first);
    for (auto& f : param_fields()) known = known || kv.first == f.name;
    if (!known) throw std::runtime_error("unknown SimulationParams field: " + kv.first);
  }
  const Node* pull = m.find("pull_fluid_to");
  if (pull && !pull->is_null()) {
    if (pull->kind != Node::Seq || pull->seq.size() < 2) throw std::runtime_error("pull_fluid_to: expected a vector");
    p.has_pull_fluid_to = 1;
    for (size_t k = 0; k < 3 && k < pull->seq.size(); k++) p.pull_fluid_to[k] = pull->seq[k].as_double();
  }
  const Node* stash = m.find("fill_stash_with");
  p.fill_stash_with = ASPH_STASH_NONE;
  if (stash && !stash->is_null()) {
    const std::string s = stash->as_string();
    if (s == "SurfaceDistanceFirstIteration") p.fill_stash_with = ASPH_STASH_SURFACE_DISTANCE_FIRST_ITERATION;
    else if (s == "SurfaceDistanceMiddle") p.fill_stash_with = ASPH_STASH_SURFACE_DISTANCE_MIDDLE;
    else throw std::runtime_error("fill_stash_with: unknown variant `" + s + "`");
  }
  const Node* diag = m.find("operator_discretization_for_diagonal");
  p.operator_discretization_for_diagonal = -1;
  if (diag && !diag->is_null()) p.operator_discretization_for_diagonal = enum_index("operator_discretization", *diag);
  return p;
}

// ------------------------------------------------------------------------------------------------ SceneConfig
struct FluidBlock { float pos[2], size[2], spacing, volume_fill_ratio, velocity[2]; };
struct SceneConfig {
  std::string boundary_type;
  float width = 0, height = 0;
  std::vector<FluidBlock> blocks;
};
inline SceneConfig scene_from_yaml(const Node& m) {
  SceneConfig s;
  const Node& b = m.at("boundary");
  s.boundary_type = b.at("type").as_string();
  s.width = float(b.at("width").as_double());
  s.height = float(b.at("height").as_double());
  const Node& blocks = m.at("blocks");
  if (blocks.kind != Node::Seq && !blocks.is_null()) throw std::runtime_error("scene: `blocks` is not a list");
  for (auto& blk : blocks.seq) {
    FluidBlock f;
    auto vec2 = [&](const char* key, float* out) {
      const Node& v = blk.at(key);
      if (v.kind != Node::Seq || v.seq.size() != 2) throw std::runtime_error(std::string("scene block: `") + key + "` is not a 2-vector");
      out[0] = float(v.seq[0].as_double()); out[1] = float(v.seq[1].as_double());
    };
    vec2("pos", f.pos); vec2("size", f.size); vec2("velocity", f.velocity);
    f.spacing = float(blk.at("spacing").as_double());
    f.volume_fill_ratio = float(blk.at("volume_fill_ratio").as_double());
    s.blocks.push_back(f);
  }
  return s;
}

struct Particles { std::vector<float> pos, vel, mass; size_t n() const { return mass.size(); } };

// add_fluid_block, simulation.rs:2915-2983: lattice of floor(box / spacing) points per axis, x-major, mass = spacing^2 * fill
inline void add_fluid_block(const FluidBlock& b, Particles& out) {
  const float spacing = b.spacing;
  const float mn[2] = {b.pos[0], b.pos[1]};
  const float mx[2] = {b.pos[0] + b.size[0], b.pos[1] + b.size[1]};
  const float particle_volume = (spacing * spacing) * b.volume_fill_ratio;
  const float particle_mass = particle_volume * 1.0f;  // INIT_REST_DENSITY
  const float box[2] = {mx[0] - mn[0], mx[1] - mn[1]};
  const long nx = long(std::floor(box[0] / spacing)), ny = long(std::floor(box[1] / spacing));
  for (long ix = 0; ix < nx; ix++) {
    float x = float(ix) * spacing;
    x = x + mn[0];
    for (long iy = 0; iy < ny; iy++) {
      float y = float(iy) * spacing;
      y = y + mn[1];
      out.pos.push_back(x); out.pos.push_back(y);
      out.vel.push_back(b.velocity[0]); out.vel.push_back(b.velocity[1]);
      out.mass.push_back(particle_mass);
    }
  }
}
inline Particles scene_particles(const SceneConfig& s) {
  Particles p;
  for (auto& b : s.blocks) add_fluid_block(b, p);
  return p;
}

// boundary handler set-up of init_fluid_sim, simulation.rs:3137-3213
inline asph_boundary scene_boundary(const SceneConfig& s, int init_boundary_handler) {
  asph_boundary b;
  std::memset(&b, 0, sizeof(b));
  const float hx = s.width / 2.0f, hy = s.height / 2.0f;
  const float mn[2] = {0.0f - hx, 0.0f - hy}, mx[2] = {0.0f + hx, 0.0f + hy};
  switch (init_boundary_handler) {
    case ASPH_BOUNDARY_ANALYTIC_OVERESTIMATE: {  // SdfPlane::new_boundary_box, sdf/sdf_plane.rs:13-20
      b.kind = ASPH_BND_PLANES; b.n_planes = 4;
      const float planes[4][3] = {{1.f, 0.f, -mn[0]}, {-1.f, 0.f, mx[0]}, {0.f, 1.f, -mn[1]}, {0.f, -1.f, mx[1]}};
      std::memcpy(b.planes, planes, sizeof(planes));
      break;
    }
    case ASPH_BOUNDARY_ANALYTIC_UNDERESTIMATE: {  // Sdf2D::new_boundary_box, sdf/sdf2d.rs:153-164
      b.kind = ASPH_BND_POLYGON; b.n_poly = 4;
      const float pts[4][2] = {{mn[0], mn[1]}, {mx[0], mn[1]}, {mx[0], mx[1]}, {mn[0], mx[1]}};
      std::memcpy(b.poly, pts, sizeof(pts));
      break;
    }
    case ASPH_BOUNDARY_NONE: b.kind = ASPH_BND_NONE; break;
    default:
      throw std::runtime_error("init_boundary_handler: Particles is out of scope (unusable in the adaptive build, particle_boundary_handler.rs:95-98)");
  }
  return b;
}

// init_simulation_params, simulation.rs:3233-3256 (adaptive build: params.h is unused, forced to 0)
inline void init_simulation_params(asph_params& p, const SceneConfig&) { p.h = 0.0; }

// ------------------------------------------------------------------------------------------------ SplitPatterns
struct SplitPatterns {
  std::vector<int32_t> offset;
  std::vector<float> pos;
  asph_split_patterns c;
};
inline void split_patterns_from_yaml(const Node& list, SplitPatterns& sp) {
  if (list.kind != Node::Seq) throw std::runtime_error("split patterns: expected a list");
  int o = 0;
  for (size_t i = 0; i < list.seq.size(); i++) {
    const Node& pts = list.seq[i].at("pos_s");
    if (pts.kind != Node::Seq || pts.seq.size() != i + 2)  // SplitPatterns::new, splitting.rs:102-108
      throw std::runtime_error("split pattern " + std::to_string(i) + " does not have " + std::to_string(i + 2) + " points");
    sp.offset.push_back(o);
    for (auto& pt : pts.seq) {
      if (pt.kind != Node::Seq || pt.seq.size() != 2) throw std::runtime_error("split pattern point is not a 2-vector");
      sp.pos.push_back(float(pt.seq[0].as_double())); sp.pos.push_back(float(pt.seq[1].as_double()));
    }
    o += int(pts.seq.size());
  }
  sp.c.max_children = int32_t(list.seq.size()) + 1;  // get_max_num_children, splitting.rs:117-119
  sp.c.offset = sp.offset.data();
  sp.c.pos_xy = sp.pos.data();
}

// ------------------------------------------------------------------------------------------------ the library
struct Library {
  void* handle = nullptr;
  decltype(&asph_create) create = nullptr;
  decltype(&asph_destroy) destroy = nullptr;
  decltype(&asph_step) step = nullptr;
  decltype(&asph_step_physics) step_physics = nullptr;
  decltype(&asph_step_adaptivity) step_adaptivity = nullptr;
  decltype(&asph_num_particles) num_particles = nullptr;
  decltype(&asph_time) time = nullptr;
  decltype(&asph_get_field) get_field = nullptr;
  decltype(&asph_get_step_info) get_step_info = nullptr;
  decltype(&asph_get_counters) get_counters = nullptr;
  decltype(&asph_last_error) last_error = nullptr;
  decltype(&asph_backend_name) backend_name = nullptr;
  // multi-GPU (bound on demand: `run --gpus N`)
  decltype(&asph_comm_unique_id) comm_unique_id = nullptr;
  decltype(&asph_create_distributed) create_distributed = nullptr;

  explicit Library(const std::string& path) {
    handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!handle) throw std::runtime_error(std::string("cannot load ") + path + ": " + dlerror() + " (there is no CPU fallback: build the CUDA library first)");
    auto sym = [&](const char* name) {
      void* s = dlsym(handle, name);
      if (!s) throw std::runtime_error(std::string("symbol missing from ") + path + ": " + name);
      return s;
    };
#define ASPH_BIND(member, name) member = reinterpret_cast<decltype(member)>(sym(#name))
    ASPH_BIND(create, asph_create); ASPH_BIND(destroy, asph_destroy); ASPH_BIND(step, asph_step);
    ASPH_BIND(step_physics, asph_step_physics); ASPH_BIND(step_adaptivity, asph_step_adaptivity);
    ASPH_BIND(num_particles, asph_num_particles); ASPH_BIND(time, asph_time); ASPH_BIND(get_field, asph_get_field);
    ASPH_BIND(get_step_info, asph_get_step_info); ASPH_BIND(get_counters, asph_get_counters);
    ASPH_BIND(last_error, asph_last_error); ASPH_BIND(backend_name, asph_backend_name);
    ASPH_BIND(comm_unique_id, asph_comm_unique_id); ASPH_BIND(create_distributed, asph_create_distributed);
#undef ASPH_BIND
  }
  Library(const Library&) = delete;
  ~Library() {}  // never dlclose: the CUDA runtime (and OpenMP in the test oracle) keep threads and atexit handlers in the library
};

inline const char* status_name(int rc) {
  static const char* names[] = {"OK", "INVALID", "UNSUPPORTED", "NONFINITE", "NEG_AII", "DENSITY", "MASS_CONSERVATION",
                                "NEIGHBOR_OVERFLOW", "CUDA", "NCCL", "CAPACITY", "NO_DEVICE"};
  return rc >= 0 && rc < 12 ? names[rc] : "?";
}

// FluidSimulation: the handle plus the calls the front ends make (simulation.rs:471-537, 1973-2796)
struct FluidSimulation {
  Library& lib;
  asph_sim* sim = nullptr;
  FluidSimulation(Library& l, const asph_params& params, const Particles& p, const asph_boundary& boundary, const SplitPatterns* split,
                  bool counters_enabled)
      : lib(l) {
    const int rc = lib.create(&params, p.pos.data(), p.vel.data(), p.mass.data(), p.n(), &boundary, split ? &split->c : nullptr,
                              counters_enabled ? 1 : 0, 0, &sim);
    if (rc != ASPH_OK) throw std::runtime_error(std::string("asph_create failed: ") + status_name(rc));
  }
  // one rank of a multi-GPU run (`run --gpus N`): this rank's share of the particles and their reference indices;
  // the library migrates them to their owner slabs at the first step
  FluidSimulation(Library& l, const asph_params& params, const Particles& p, const std::vector<uint32_t>& global_index, uint64_t n_global,
                  const asph_boundary& boundary, const SplitPatterns* split, bool counters_enabled, const uint8_t nccl_id[128], int rank, int n_ranks,
                  int device)
      : lib(l) {
    const int rc = lib.create_distributed(&params, p.pos.data(), p.vel.data(), p.mass.data(), global_index.data(), p.n(), n_global, &boundary,
                                          split ? &split->c : nullptr, counters_enabled ? 1 : 0, 0, nccl_id, rank, n_ranks, device, &sim);
    if (rc != ASPH_OK) throw std::runtime_error(std::string("asph_create_distributed failed: ") + status_name(rc));
  }
  FluidSimulation(const FluidSimulation&) = delete;
  ~FluidSimulation() { if (sim) lib.destroy(sim); }
  void check(int rc) const {
    if (rc == ASPH_OK) return;
    const char* msg = lib.last_error(sim);
    throw std::runtime_error(std::string("asph error ") + std::to_string(rc) + " (" + status_name(rc) + "): " + (msg ? msg : ""));
  }
  float single_step(const asph_params& p) { float dt = 0; check(lib.step(sim, &p, &dt)); return dt; }
  float single_step_without_adaptivity(const asph_params& p) { float dt = 0; check(lib.step_physics(sim, &p, &dt)); return dt; }
  void single_step_adaptivity(const asph_params& p, float dt) { check(lib.step_adaptivity(sim, &p, dt)); }
  uint64_t num_fluid_particles() const { return lib.num_particles(sim); }
  double time() const { return lib.time(sim); }
  asph_step_info step_info() const { asph_step_info i; std::memset(&i, 0, sizeof(i)); lib.get_step_info(sim, &i); return i; }
  std::vector<float> field(int id, int comps) {
    std::vector<float> v(size_t(num_fluid_particles()) * size_t(comps));
    check(lib.get_field(sim, id, v.data(), v.size() * sizeof(float)));
    return v;
  }
};

// ------------------------------------------------------------------------------------------------ statistics
struct StatisticsRecorder {  // ValueCounters, simulation.rs:137-157
  std::map<std::string, std::vector<double>> values;
  void add(const std::string& k, double v) { values[k].push_back(v); }
  void record_step(const asph_step_info& i) {
    add("particle-count", double(i.n_particles_begin));
    add("dt", double(i.dt));
    if (i.div_iterations > 0) add("div-iterations", i.div_iterations);
    if (i.density_iterations > 0) add("density-iterations", i.density_iterations);
  }
  double avg(const std::string& k) const {
    auto it = values.find(k);
    if (it == values.end() || it->second.empty()) return std::nan("");
    double s = 0;
    for (double v : it->second) s += v;
    return s / double(it->second.size());
  }
  // write_statistics, simulation.rs:3279-3359
  std::string write_statistics(FluidSimulation& sim) const {
    static const char* labels[ASPH_PC_COUNT] = {"simulation-step", "neighborhood", "level-estimation", "div-solver", "density-solver", "adaptivity"};
    double ms[ASPH_PC_COUNT];
    uint64_t calls[ASPH_PC_COUNT];
    sim.lib.get_counters(sim.sim, ms, calls);
    char buf[512];
    std::string s;
    const double pc = avg("particle-count");
    std::snprintf(buf, sizeof(buf), "$%.2f\\si{\\second}$ & %ld & %.2f & %.2f & - \\\\\n\nsimulation-time: %gms\n\n", ms[0] / 1000.0,
                  std::isnan(pc) ? 0L : std::lround(pc), avg("div-iterations"), avg("density-iterations"), ms[0]);
    s += buf;
    std::map<std::string, int> order;
    for (int k = 0; k < ASPH_PC_COUNT; k++) order[labels[k]] = k;
    for (auto& kv : order)
      if (calls[kv.second]) { std::snprintf(buf, sizeof(buf), "%s: avg:%gms\n", kv.first.c_str(), ms[kv.second] / double(calls[kv.second])); s += buf; }
    s += "\n";
    for (auto& kv : values) {
      double mn = kv.second[0], mx = kv.second[0], sum = 0;
      for (double v : kv.second) { mn = std::min(mn, v); mx = std::max(mx, v); sum += v; }
      std::snprintf(buf, sizeof(buf), "%s: min:%g max:%g avg:%g\n", kv.first.c_str(), mn, mx, sum / double(kv.second.size()));
      s += buf;
    }
    return s;
  }
};

}  // namespace host
